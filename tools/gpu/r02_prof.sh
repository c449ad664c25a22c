#!/bin/bash
# full GPU test suite, per-precision launch lists, default bench with all legs, and
# ncu --set full on the dominant kernels of the default mode
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$? $(grep -E 'passed|failed' gpurun_out/t_all.log | tail -1)"
grep -E "rel-L2 [0-9.e-]+  out-of-tol|^bf16 |^FAILED|^ERROR|^E  |worst" gpurun_out/t_all.log | head -70 | cut -c1-220
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); r = d['roofline']
    print('DEFAULT', d['config']['gemm_precision'], 'ms/step', round(d['ms_per_step'],3), 'M/s', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1),
          'launches/step', d['gpu_launches']//d['steps'], 'gemm ms', round(r['gemm_ms_per_step'],3), {k: round(v['ms'],3) for k, v in r['per_kernel'].items()})
    print(' issued frac', r.get('issued_mma_frac_of_pipe_peak'), 'peaks', d.get('peaks_measured_here'))
    print(' library', {k: (round(v['ms_per_step'],2) if isinstance(v, dict) and 'ms_per_step' in v else v) for k, v in d.get('library_baseline', {}).items() if k != 'what'})
    for o in d.get('other_configs', []):
        print(' other', o.get('config') if isinstance(o.get('config'), str) else o.get('config', {}).get('workload'), o.get('ms_per_step'), o.get('error'), o.get('graph_error'))
    print(' cpu', d.get('cpu_baseline', {}).get('value'), 'clocks', d.get('clocks'))
except Exception as e:
    print('default bench failed:', e); print(open('gpurun_out/bench_default.err').read()[-1500:])
PY
for p in ${PRECS:-3xf16 bf16 3xtf32}; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_$p.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras --precision $p > gpurun_out/ncu_launch_$p.log 2>&1
  echo "launch list $p rc=$?"
done
# ncu --set full: one step's GEMM launches (NT pair / NT / TN) + the three BatchNorm passes, default mode
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|bn_prelu" -s 400 -c 60 -o gpurun_out/prof_r02_3xf16 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras --precision 3xf16 > gpurun_out/ncu_full_3xf16.log 2>&1
echo "ncu full 3xf16 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|bn_prelu" -s 400 -c 60 -o gpurun_out/prof_r02_bf16 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras --precision bf16 > gpurun_out/ncu_full_bf16.log 2>&1
echo "ncu full bf16 rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -3
