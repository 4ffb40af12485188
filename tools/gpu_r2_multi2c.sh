#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 tools/gpu_verify_windows.py 400000 3 96 > gpurun_out/verify_windows_n2.log 2>&1; echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/verify_windows_n2.log | tail -20
