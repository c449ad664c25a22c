#!/bin/bash
# Round 2, first contact of the 16-bit tensor-core modes with hardware.  Each risky group runs
# in its own process (a trapped kernel poisons the CUDA context of its process only).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() {  # name, timeout, pytest args...
  local name=$1 to=$2; shift 2
  timeout $to python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $name rc=$? :: $(grep -E '^[0-9]+ (passed|failed)|passed|failed' gpurun_out/$name.log | tail -1)"
  grep -E "^FAILED|^ERROR|timed out|max err|Error:|rel-L2|mismatch" gpurun_out/$name.log | head -${SHOW:-8} | cut -c1-260
}
run k_regress 300 tests/test_tc_gemm_gpu.py -k "test_tc_gemm_nt and (1- or 0-) or test_tc_gemm_tn and (1- or 0-) or split"
run k_nt_bf16 200 tests/test_tc_gemm_gpu.py -k "test_tc_gemm_nt and 2-"
run k_nt_3xf16 200 tests/test_tc_gemm_gpu.py -k "test_tc_gemm_nt and 3-"
run k_tn_bf16 200 tests/test_tc_gemm_gpu.py -k "test_tc_gemm_tn and 2-"
run k_tn_3xf16 200 tests/test_tc_gemm_gpu.py -k "test_tc_gemm_tn and 3-"
run k_nt_out16 300 tests/test_tc_gemm_gpu.py -k "bf16_out_and_alpha_dev"
run k_tn_16shapes 200 tests/test_tc_gemm_gpu.py -k "tn_16bit_shapes"
if grep -q failed gpurun_out/k_tn_bf16.log; then
  # alternative reading of the MN-major descriptor: swap LBO / SBO
  PASE_B200_TN16_DESC="1024,8192" run k_tn_bf16_swap 200 tests/test_tc_gemm_gpu.py -k "test_tc_gemm_tn and 2-"
fi
run k_elementwise 400 tests/test_kernels_gpu.py
SHOW=20 run enc_golden 900 tests/test_encoder_gpu.py -k "golden or bf16_mode or tf32_mode"
SHOW=20 run enc_shapes 900 tests/test_encoder_gpu.py -k "benchmark_shape or full_length or bf16_benchmark" -s
run rest 600 tests/test_graph_gpu.py tests/test_heads_gpu.py tests/test_abi.py
for p in 3xtf32 3xf16 bf16 tf32; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision $p 2> gpurun_out/bench_$p.err | tail -1 > gpurun_out/bench_$p.json
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$p.json')); r = d['roofline']
    print('$p ms/step', round(d['ms_per_step'],3), 'M/s', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1),
          'gemm ms', round(r['gemm_ms_per_step'],3), {k: round(v['ms'],3) for k, v in r['per_kernel'].items()}, 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('$p bench failed:', e); print(open('gpurun_out/bench_$p.err').read()[-600:])
PY
done
