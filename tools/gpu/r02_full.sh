#!/bin/bash
# the round's closing run on one GPU: whole GPU suite, smoke, default bench with every leg,
# reference arm
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$? $(grep -E 'passed|failed' gpurun_out/t_all.log | tail -1)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/t_all.log | head -30 | cut -c1-220
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); r = d['roofline']
    print('DEFAULT', d['config']['gemm_precision'], 'ms/step', round(d['ms_per_step'],3), 'M/s', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'e2e ms', round(d['e2e']['ms_per_step'],3),
          'launches/step', d['gpu_launches']//d['steps'], 'gemm ms', round(r['gemm_ms_per_step'],3), {k: round(v['ms'],3) for k, v in r['per_kernel'].items()})
    print(' frac', round(r['frac'],3), 'issued frac', r.get('issued_mma_frac_of_pipe_peak'), 'traffic', r.get('traffic'), 'peaks', d.get('peaks_measured_here'))
    print(' library', {k: (round(v['ms_per_step'],2) if isinstance(v, dict) and 'ms_per_step' in v else v) for k, v in d.get('library_baseline', {}).items() if k != 'what'})
    for o in d.get('other_configs', []):
        print(' other', o.get('config') if isinstance(o.get('config'), str) else o.get('config', {}).get('workload'), o.get('ms_per_step'), o.get('error'), o.get('graph_error'))
    print(' cpu', d.get('cpu_baseline', {}).get('value'), 'clocks', d.get('clocks'))
except Exception as e:
    print('default bench failed:', e); print(open('gpurun_out/bench_default.err').read()[-1500:])
PY
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
