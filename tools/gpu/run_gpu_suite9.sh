#!/bin/bash
# full gpu tests, default bench, ncu launch list, and an A/B of the epilogue poll back-off
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$?"; grep -E "passed|failed|^FAILED|^E  .*Error" gpurun_out/t_all.log | head -12 | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1
tail -n 1 gpurun_out/bench_default.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('ms/step', round(d['ms_per_step'],3), 'value M/s', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'launches', d['gpu_launches'], 'gemm ms', round(r['gemm_ms_per_step'],3), 'TF', round(r['achieved'],1))"
for ns in 0 200 1000; do
PASE_B200_EPI_SLEEP_NS=$ns timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('epi_sleep $ns ms/step', round(d['ms_per_step'],3), 'gemm ms', round(r['gemm_ms_per_step'],3), {k:round(v['ms'],3) for k,v in r['per_kernel'].items()})"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_3xtf32.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu rc=$?"
