#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_targets.py tests/test_host_emulated.py -q -m gpu -p no:cacheprovider > gpurun_out/t_lps.log 2>&1
echo "targets gpu tests rc=$? $(grep -E 'passed|failed' gpurun_out/t_lps.log | tail -1)"; grep -E "^FAILED|^E  " gpurun_out/t_lps.log | head -10 | cut -c1-300
timeout 300 python - <<'PY'
import torch, json, bench
dev = torch.device("cuda", 0)
print(json.dumps(bench.time_lps_targets(dev)))
PY
