#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python tools/gpu_verify_windows.py 400000 3 96 > gpurun_out/verify_windows_n1.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/verify_windows_n1.log
