#!/bin/bash
# launch lists (every kernel's device time) of one step per precision, then ncu --set full on
# the dominant kernels of the default mode
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -m gpu -p no:cacheprovider -s -k "benchmark_shape or bf16" > gpurun_out/enc_shapes2.log 2>&1
echo "enc_shapes2 rc=$? $(grep -E 'passed|failed' gpurun_out/enc_shapes2.log | tail -1)"
grep -E "rel-L2|^bf16|FAILED|^E  " gpurun_out/enc_shapes2.log | head -60 | cut -c1-200
for p in ${PRECS:-bf16 3xf16 3xtf32}; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_$p.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --precision $p > gpurun_out/ncu_launch_$p.log 2>&1
  echo "launch list $p rc=$?"
done
