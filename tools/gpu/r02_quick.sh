#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py tests/test_encoder_gpu.py -q -m gpu -p no:cacheprovider -k "tc_gemm or golden or benchmark_shape_against" > gpurun_out/q_tests.log 2>&1
echo "tests rc=$? $(grep -E 'passed|failed' gpurun_out/q_tests.log | tail -1)"; grep -E "^FAILED|^E  " gpurun_out/q_tests.log | head -10 | cut -c1-220
for p in ${PRECS:-3xf16 bf16 3xtf32}; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --precision $p 2> gpurun_out/q_$p.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$p ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'gemm ms', round(r['gemm_ms_per_step'],3), {k: round(v['ms'],3) for k, v in r['per_kernel'].items()}, 'traffic', r['traffic'])"
done
timeout 300 python bench.py --workload workers --steps 5 --warmup 3 2> gpurun_out/q_workers.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('workers ms/step', round(d['ms_per_step'],3), 'graph', d['cuda_graph'], d['graph_error'])"
