#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python tools/gpu/debug_4001.py 2>&1 | tail -22
PASE_B200_BN_STREAM=0 timeout 300 python tools/gpu/debug_4001.py 2>&1 | tail -22
PASE_B200_BN_DU=1 timeout 300 python tools/gpu/debug_4001.py 2>&1 | tail -22
