#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gemm_gpu.py -q -m gpu -x > gpurun_out/t_tc.log 2>&1
echo "tc rc=$?"; tail -n 30 gpurun_out/t_tc.log
timeout 900 python -m pytest tests/test_tc_gemm_gpu.py -q -m gpu > gpurun_out/t_tc_all.log 2>&1
echo "tc-all rc=$?"; tail -n 40 gpurun_out/t_tc_all.log | cut -c1-300
timeout 900 python -m pytest tests/test_heads_gpu.py tests/test_encoder_gpu.py -q -m gpu > gpurun_out/t_heads.log 2>&1
echo "heads+enc rc=$?"; tail -n 30 gpurun_out/t_heads.log | cut -c1-300
# ncu: launch list of two steps, then a full-set capture of the two GEMM kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_simt.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu-launch rc=$?"; tail -n 3 gpurun_out/ncu_launch.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 12 -c 2 -o gpurun_out/prof_r01_gemm_nt python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu-full rc=$?"; tail -n 3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
