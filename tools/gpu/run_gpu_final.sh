#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1
echo "gpu tests rc=$?"; grep -E "passed|failed" gpurun_out/t_all.log | tail -1
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.log 2>&1
tail -n 1 gpurun_out/bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('ms/step', round(d['ms_per_step'],3), 'value M/s', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'gemm ms', round(r['gemm_ms_per_step'],3), 'TF', round(r['achieved'],1), 'frac', round(r['frac'],4))"
