#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/t_nccl.log 2>&1
echo "dp tests rc=$? $(grep -E 'passed|failed|skipped' gpurun_out/t_nccl.log | tail -1)"; grep -E "worst|AssertionError" gpurun_out/t_nccl.log | head -12 | cut -c1-600
