#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$?"; grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/t_all.log | head -12 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1
grep -v "^{" gpurun_out/bench_default.log | tail -3 | cut -c1-200
tail -n 1 gpurun_out/bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('ms/step', round(d['ms_per_step'],3), 'value M/s', round(d['value']/1e6,1), 'e2e', d['e2e'], 'launches', d['gpu_launches'], 'gemm ms', round(r['gemm_ms_per_step'],3), 'TF', round(r['achieved'],1), 'cpu', d.get('cpu_baseline',{}).get('value'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_3xtf32.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu rc=$?"
