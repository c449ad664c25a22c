#!/bin/bash
# 4 GPUs: the fused peer-memory DP update at N=4 (and NCCL for comparison)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'M/s', round(d['value']/1e6,1), 'e2e ms', round(d['e2e']['ms_per_step'],3), (d['config'].get('grad_allreduce') or '')[:60])
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open(sys.argv[2].replace('.json','.err')).read()[-1500:])
PY
}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/n4.json 2> gpurun_out/n4.err; show n4_peer gpurun_out/n4.json; grep -i "unavailable\|error" gpurun_out/n4.err | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-extras --dp nccl > gpurun_out/n4b.json 2> gpurun_out/n4b.err; show n4_nccl gpurun_out/n4b.json
