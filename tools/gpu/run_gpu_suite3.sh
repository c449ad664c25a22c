#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py tests/test_kernels_gpu.py -q -m gpu > gpurun_out/t_tc.log 2>&1
echo "tc+kernels rc=$?"; grep -E "passed|failed|^FAILED|AssertionError" gpurun_out/t_tc.log | head -30 | cut -c1-250
timeout 120 python tools/tf32_probe.py 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_heads_gpu.py tests/test_encoder_gpu.py -q -m gpu > gpurun_out/t_enc.log 2>&1
echo "heads+enc rc=$?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/t_enc.log | head -40 | cut -c1-300
for p in 3xtf32 tf32; do
  timeout 600 python bench.py --steps 10 --warmup 3 --precision $p --no-cpu-baseline > gpurun_out/bench_$p.log 2>&1
  echo "bench $p rc=$?"; tail -n 1 gpurun_out/bench_$p.log | cut -c1-1500
done
