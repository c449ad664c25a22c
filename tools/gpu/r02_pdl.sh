#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in nt trig off; do
unset PASE_B200_LIB; export PASE_B200_PDL=1
[ $v = nt ] && export PASE_B200_LIB=$PWD/pase_b200/csrc/libpase_b200_nt.so
[ $v = off ] && export PASE_B200_PDL=0
for p in ${PRECS:-3xf16 bf16}; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --precision $p 2> gpurun_out/q_$p.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('pdl $v $p ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'gemm ms', round(r['gemm_ms_per_step'],3))"
done
done
