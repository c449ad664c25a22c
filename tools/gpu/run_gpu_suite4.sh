#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -q -m gpu > gpurun_out/t_tc.log 2>&1
echo "tc rc=$?"; grep -E "passed|failed|^FAILED|AssertionError|timed out" gpurun_out/t_tc.log | head -30 | cut -c1-250
timeout 1500 python -m pytest tests/test_heads_gpu.py tests/test_encoder_gpu.py -q -m gpu > gpurun_out/t_enc.log 2>&1
echo "heads+enc rc=$?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/t_enc.log | head -40 | cut -c1-300
for p in 3xtf32 tf32; do
  timeout 600 python bench.py --steps 10 --warmup 3 --precision $p --no-cpu-baseline > gpurun_out/bench_$p.log 2>&1
  echo "bench $p rc=$?"; tail -n 1 gpurun_out/bench_$p.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], 'gemm ms', round(r['gemm_ms_per_step'],3), 'TF', round(r['achieved'],1), {k:(v['launches'], round(v['ms'],3), round(v['tflops'],1)) for k,v in r['per_kernel'].items()})
"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_3xtf32.csv python bench.py --steps 1 --warmup 3 --precision 3xtf32 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu-launch rc=$?"
