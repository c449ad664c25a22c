#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
python tools/gpu/debug_head.py 3xf16 2>&1 | tail -12
python tools/gpu/debug_head.py 3xtf32 2>&1 | tail -12
PASE_B200_TN_2CTA=0 python tools/gpu/debug_head.py 3xf16 2>&1 | tail -12
