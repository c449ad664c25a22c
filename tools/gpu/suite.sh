#!/bin/bash
# full gpu tests, default bench, tf32 / workers operating points, ncu launch list
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$?"; grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/t_all.log | head -12 | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1
tail -n 1 gpurun_out/bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('ms/step', round(d['ms_per_step'],3), 'value M/s', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'launches', d['gpu_launches'], 'gemm ms', round(r['gemm_ms_per_step'],3), 'TF', round(r['achieved'],1))"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision tf32 2>/dev/null | tail -1 > gpurun_out/bench_tf32.json
python -c "
import json; d = json.load(open('gpurun_out/bench_tf32.json')); print('tf32 ms/step', round(d['ms_per_step'],3), 'value M/s', round(d['value']/1e6,1))"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload workers 2>/dev/null | tail -1 > gpurun_out/bench_workers.json
python -c "
import json; d = json.load(open('gpurun_out/bench_workers.json')); print('workers ms/step', round(d['ms_per_step'],3), 'value M/s', round(d['value']/1e6,1))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_3xtf32.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu rc=$?"
