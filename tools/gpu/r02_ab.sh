#!/bin/bash
# correctness of the window / TN-pair kernels, then A/B of each lever on the step time
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() {
  local name=$1 to=$2; shift 2
  timeout $to python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $name rc=$? :: $(grep -E 'passed|failed' gpurun_out/$name.log | tail -1)"
  grep -E "^FAILED|^ERROR|timed out|max err|Error:|mismatch" gpurun_out/$name.log | head -${SHOW:-8} | cut -c1-240
}
run ab_gemm 400 tests/test_tc_gemm_gpu.py
run ab_kern 400 tests/test_kernels_gpu.py
PASE_B200_BN_RUN=8 run ab_kern8 400 tests/test_kernels_gpu.py
SHOW=30 run ab_enc 900 tests/test_encoder_gpu.py -s
run ab_heads 900 tests/test_heads_gpu.py tests/test_flat_adam.py tests/test_graph_gpu.py
bench() {  # label, env...
  local label=$1; shift
  for p in 3xf16 bf16; do
    env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --precision $p 2> gpurun_out/ab_$label_$p.err | tail -1 > gpurun_out/ab_${label}_$p.json
    python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_${label}_$p.json')); r = d['roofline']
    print('$label $p ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'gemm ms', round(r['gemm_ms_per_step'],3), {k: round(v['ms'],3) for k, v in r['per_kernel'].items()})
except Exception as e:
    print('$label $p failed', e)
PY
  done
}
bench run4 X=1
bench run8 PASE_B200_BN_RUN=8
