#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
# launch 0 of the step = sinc fwd; launches 5.. = mid layers.  warmup 3 + 1 step: skip the first 3 steps' launches (19 nt each)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_nt_kernel -s 57 -c 7 -o gpurun_out/prof_r01_tc_nt python bench.py --steps 1 --warmup 3 --precision 3xtf32 --no-cpu-baseline > gpurun_out/ncu_full_tc.log 2>&1
echo "ncu rc=$?"; tail -n 4 gpurun_out/ncu_full_tc.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
