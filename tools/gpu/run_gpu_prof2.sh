#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 87 -c 29 -o gpurun_out/prof_r01_tc_gemm_v3 python bench.py --steps 1 --warmup 3 --precision 3xtf32 --no-cpu-baseline > gpurun_out/ncu_full_tc.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_full_tc.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_3xtf32.csv python bench.py --steps 1 --warmup 3 --precision 3xtf32 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1; tail -n 1 gpurun_out/bench_default.log | cut -c1-400
