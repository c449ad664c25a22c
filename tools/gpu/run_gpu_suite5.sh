#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$?"; grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/t_all.log | head -12 | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.log 2>&1
tail -n 1 gpurun_out/bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], 'gemm ms', round(r['gemm_ms_per_step'],3), 'TF', round(r['achieved'],1), {k:(v['launches'], round(v['ms'],3), round(v['tflops'],1)) for k,v in r['per_kernel'].items()})"
