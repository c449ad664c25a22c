#!/bin/bash
# GPU-box suite: kernel tests, encoder parity, smoke, short bench.  Logs -> gpurun_out/
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu > gpurun_out/t_kernels.log 2>&1
echo "kernels rc=$?"; tail -n 25 gpurun_out/t_kernels.log
timeout 1500 python -m pytest tests/test_encoder_gpu.py -q -m gpu > gpurun_out/t_encoder.log 2>&1
echo "encoder rc=$?"; tail -n 40 gpurun_out/t_encoder.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -n 3 gpurun_out/bench.log
