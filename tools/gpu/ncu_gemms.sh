#!/bin/bash
# ncu --set full on one step's worth of tcgen05 GEMM launches (NT: 10 launches, TN: 6)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_nt -s 57 -c 10 -o gpurun_out/prof_r01_nt_v4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_nt_v4.log 2>&1
echo "nt rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_tn -s 30 -c 6 -o gpurun_out/prof_r01_tn_v4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_tn_v4.log 2>&1
echo "tn rc=$?"
