#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for w in 0 1; do
  PASE_B200_TC_WINDOW=$w timeout 600 ncu --set full --clock-control none -k regex:tc_gemm_nt -s 62 -c 1 -o gpurun_out/prof_blk5_w$w python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_blk5_w$w.log 2>&1
  echo "w=$w rc=$?"
done
