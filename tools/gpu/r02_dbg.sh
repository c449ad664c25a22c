#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python tools/gpu/kineto_step.py 3xf16 graph > gpurun_out/kineto_3xf16.md 2> gpurun_out/kineto_err.log; echo rc=$?
tail -5 gpurun_out/kineto_err.log
head -30 gpurun_out/kineto_3xf16.md
