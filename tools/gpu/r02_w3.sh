#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_heads_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/w3_tests.log 2>&1
echo "heads rc=$? $(grep -E 'passed|failed' gpurun_out/w3_tests.log | tail -1)"; grep -E "^FAILED|^E  " gpurun_out/w3_tests.log | head -12 | cut -c1-240
for f in "" "--no-fuse-heads"; do
timeout 300 python bench.py --workload workers --steps 5 --warmup 3 $f 2> gpurun_out/w3_workers.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('workers $f ms/step', round(d['ms_per_step'],3), 'graph', d['cuda_graph'], d['graph_error'], 'loss', d['loss'], d['config'].get('fused_regression_heads'))" || tail -5 gpurun_out/w3_workers.err
done
