#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/t_nccl.log 2>&1
echo "dp tests rc=$? $(grep -E 'passed|failed|skipped' gpurun_out/t_nccl.log | tail -1)"; grep -E "worst|AssertionError" gpurun_out/t_nccl.log | head -12 | cut -c1-600
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'M/s', round(d['value']/1e6,1), 'e2e ms', round(d['e2e']['ms_per_step'],3), (d['config'].get('grad_allreduce') or '')[:60])
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open(sys.argv[2].replace('.json','.err')).read()[-1500:])
PY
}
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/n1.json 2> gpurun_out/n1.err; show n1 gpurun_out/n1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/n2.json 2> gpurun_out/n2.err; show n2_peer gpurun_out/n2.json; grep -i "unavailable\|error" gpurun_out/n2.err | head -3
