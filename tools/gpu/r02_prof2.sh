#!/bin/bash
# launch lists per precision + ncu --set full of one step's GEMM / BatchNorm launches, exported
# to CSV on the box (the .ncu-rep files stay in /tmp: too large to bring back)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for p in ${PRECS:-3xf16 bf16 3xtf32}; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_$p.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras --precision $p > gpurun_out/ncu_launch_$p.log 2>&1
  echo "launch list $p rc=$? $(wc -c < gpurun_out/launches_r02_$p.csv) bytes"
done
for p in ${FULLS:-3xf16 bf16}; do
  timeout 900 ncu --set full --clock-control none -k regex:"tc_gemm|bn_prelu|bn_bwd_stream" -s 424 -c 53 -o /tmp/prof_r02_$p python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras --precision $p > gpurun_out/ncu_full_$p.log 2>&1
  echo "ncu full $p rc=$?"
  ncu -i /tmp/prof_r02_$p.ncu-rep --page raw --csv > gpurun_out/ncu_raw_r02_$p.csv 2>/dev/null
  echo "raw csv $p: $(wc -c < gpurun_out/ncu_raw_r02_$p.csv) bytes"
done
du -sh gpurun_out
for p in 3xf16 bf16; do
  timeout 300 python tools/gpu/kineto_step.py $p graph > gpurun_out/kineto_$p.md 2> gpurun_out/kineto_err.log
  grep "kernels, span" gpurun_out/kineto_$p.md
done
timeout 300 python tools/gpu/kineto_workers.py 3xf16 > gpurun_out/kineto_workers.md 2> gpurun_out/kineto_w_err.log; head -3 gpurun_out/kineto_workers.md | tail -1
