#!/bin/bash
# default-path sanity, then an A/B of the CTA-pair NT kernel (PASE_B200_TC_2CTA=0/1)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1
echo "default gpu tests rc=$?"; grep -E "passed|failed" gpurun_out/t_all.log | tail -1
PASE_B200_TC_2CTA=0 timeout 150 python -m pytest tests/test_tc_gemm_gpu.py -q -m gpu -x > gpurun_out/t_2cta.log 2>&1
rc=$?; echo "1-CTA (PASE_B200_TC_2CTA=0) gemm tests rc=$rc"; grep -E "passed|failed|Error|timed out|max err|mismatch" gpurun_out/t_2cta.log | head -8 | cut -c1-250
if [ $rc -eq 0 ]; then
  PASE_B200_TC_2CTA=0 timeout 200 python -m pytest tests/test_encoder_gpu.py -q -m gpu -x > gpurun_out/t_2cta_enc.log 2>&1
  echo "1-CTA encoder tests rc=$?"; grep -E "passed|failed" gpurun_out/t_2cta_enc.log | tail -1
  for v in 0 1; do
    PASE_B200_TC_2CTA=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('2cta=$v ms/step', round(d['ms_per_step'],3), 'gemm ms', round(r['gemm_ms_per_step'],3), {k:round(v['ms'],3) for k,v in r['per_kernel'].items()})"
  done
fi
