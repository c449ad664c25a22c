#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -q -m gpu > gpurun_out/t_tc.log 2>&1
echo "tc rc=$?"; grep -E "passed|failed|^FAILED|AssertionError|timed out" gpurun_out/t_tc.log | head -12 | cut -c1-250
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_heads_gpu.py -q -m gpu > gpurun_out/t_enc.log 2>&1
echo "enc rc=$?"; grep -E "passed|failed|^FAILED" gpurun_out/t_enc.log | head -8 | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_eager.log 2>&1
tail -n 1 gpurun_out/bench_eager.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('ms/step', round(d['ms_per_step'],3), 'gemm ms', round(r['gemm_ms_per_step'],3), 'TF', round(r['achieved'],1), {k:(v['launches'], round(v['ms'],3), round(v['tflops'],1)) for k,v in r['per_kernel'].items()})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tmp.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
