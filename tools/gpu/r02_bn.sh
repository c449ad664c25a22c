#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_encoder_gpu.py -q -m gpu -p no:cacheprovider -k "bn_prelu or golden or benchmark_shape_against or bf16" > gpurun_out/q_tests.log 2>&1
echo "tests rc=$? $(grep -E 'passed|failed' gpurun_out/q_tests.log | tail -1)"; grep -E "^FAILED|^E  " gpurun_out/q_tests.log | head -10 | cut -c1-220
for tile in 4096 2048; do
export PASE_B200_BN_TILE=$tile
for p in ${PRECS:-3xf16 bf16}; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --precision $p 2> gpurun_out/q_$p.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('tile $tile $p ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'gemm ms', round(r['gemm_ms_per_step'],3))"
done
timeout 600 python tools/gpu/kineto_step.py 3xf16 graph > gpurun_out/kineto_3xf16_$tile.md 2> gpurun_out/kineto_err.log
grep "bn_\|kernels, span" gpurun_out/kineto_3xf16_$tile.md | head -4
grep "bn_bwd_stream" gpurun_out/kineto_3xf16_$tile.md | tail -6
done
