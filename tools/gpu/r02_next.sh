#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_heads_gpu.py tests/test_encoder_gpu.py -k "not benchmark_shape" -q -m gpu -p no:cacheprovider > gpurun_out/q_tests.log 2>&1
echo "tests rc=$? $(grep -E 'passed|failed' gpurun_out/q_tests.log | tail -1)"; grep -E "^FAILED|^E  " gpurun_out/q_tests.log | head -10 | cut -c1-220
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --precision 3xf16 2> gpurun_out/q_3xf16.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('3xf16 ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'gemm ms', round(r['gemm_ms_per_step'],3))"
timeout 600 python tools/gpu/kineto_step.py 3xf16 graph > gpurun_out/kineto_3xf16.md 2> gpurun_out/kineto_err.log
grep "kernels, span\|colsum\|absmax" gpurun_out/kineto_3xf16.md | head -8
timeout 600 python tools/gpu/kineto_workers.py 3xf16 > gpurun_out/kineto_workers.md 2> gpurun_out/kineto_w_err.log; head -1 gpurun_out/kineto_workers.md; grep "colsum\|absmax" gpurun_out/kineto_workers.md | head -3
