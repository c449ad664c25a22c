#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for f in "" "--no-prefetch"; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras $f 2> gpurun_out/w_e2e.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench $f ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['e2e'].get('input_prefetch'))"
done
timeout 300 python bench.py --workload workers --steps 5 --warmup 3 2> gpurun_out/w_workers.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('workers ms/step', round(d['ms_per_step'],3), 'graph', d['cuda_graph'], d['graph_error'], 'loss', d['loss'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_workers.csv python bench.py --workload workers --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_launch_workers.log 2>&1
echo "workers launch list rc=$? $(wc -c < gpurun_out/launches_r02_workers.csv)"
