#!/bin/bash
# Round 2, lease 28 (last GPU minutes): the one-CTA-per-SM marching kernel on the marching parity tests and on the 2,016-step call
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_march1b.txt
nvidia-smi -L > $B 2>&1
export TROUTE_B200_LIB=$PWD/t-route_b200/troute_b200/lib/variants/libtroute_b200_march1.so
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "marching or network_bits" > gpurun_out/pytest_march1.log 2>&1; echo "pytest (variant) rc=$? $(tail -1 gpurun_out/pytest_march1.log)" >> $B
timeout 300 python bench.py --workload conus-lp7d --windows 1 --nsteps 2016 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-trip-order --verify-segments 20000 > gpurun_out/ab_march1_onecall.json 2> gpurun_out/ab_march1_onecall.err
echo "march1_onecall rc=$? $(python tools/ab_line.py gpurun_out/ab_march1_onecall.json)" >> $B
cat $B
