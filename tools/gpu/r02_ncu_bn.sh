#!/bin/bash
# ncu --set full of one step's BatchNorm-backward launches (staged kernels); CSV export on the box
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_stream|bn_prelu_pad_fwd" -s 72 -c 24 -o /tmp/prof_bn python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras --precision 3xf16 > gpurun_out/ncu_bn.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_bn.ncu-rep --page raw --csv > gpurun_out/ncu_raw_bn.csv 2>/dev/null
echo "raw csv: $(wc -c < gpurun_out/ncu_raw_bn.csv) bytes"
ncu -i /tmp/prof_bn.ncu-rep --page details --csv > gpurun_out/ncu_details_bn.csv 2>/dev/null
echo "details csv: $(wc -c < gpurun_out/ncu_details_bn.csv) bytes"
