#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -rA > gpurun_out/pytest_gpu_multi_n2d.log 2>&1; echo "pytest multi rc=$? $(tail -1 gpurun_out/pytest_gpu_multi_n2d.log)"; grep "^OK\|Error\|assert" gpurun_out/pytest_gpu_multi_n2d.log | head -20
bash tools/gpu_r2_multi2c.sh
