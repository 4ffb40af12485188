#!/bin/bash
# Round 2, lease 21: the Python mirror after the host-side work (one trt_route call, compact reservoir inflows, cache keys):
# its GPU tests, then its end-to-end cost per call on the bench network.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_mirror.txt
{ nproc; nvidia-smi -L; free -g | head -2; } > $B 2>&1
timeout 1800 python -m pytest tests/test_gpu_api.py tests/test_lowercolorado.py tests/test_lowercolorado_lakes.py tests/test_route_windows.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/pytest_gpu_mirror.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_mirror.log)" >> $B
timeout 1500 python tools/mirror_overhead.py 2729077 288 gpu > gpurun_out/mirror_overhead_gpu.txt 2>&1; echo "mirror rc=$?" >> $B
cat gpurun_out/mirror_overhead_gpu.txt >> $B
cat $B
