#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python tools/gpu/debug_fused2.py 1 60 2>&1 | tail -12
timeout 600 python tools/gpu/debug_fused2.py 0 60 2>&1 | tail -12
