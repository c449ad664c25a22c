#!/bin/bash
# 2 GPUs: DP correctness tests (NCCL flat all-reduce + pipelined step; fused peer-memory update),
# then the N=2 bench (peer-memory fused update vs NCCL all-reduce) next to N=1 on the same box
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -5
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/t_nccl.log 2>&1
echo "dp tests rc=$? $(grep -E 'passed|failed|skipped' gpurun_out/t_nccl.log | tail -1)"; grep -E "worst|Error|^E " gpurun_out/t_nccl.log | head -12 | cut -c1-240
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
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras --dp nccl > gpurun_out/n2b.json 2> gpurun_out/n2b.err; show n2_nccl gpurun_out/n2b.json
