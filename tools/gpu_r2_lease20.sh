#!/bin/bash
# Round 2, lease 20: marching lanes in groups of four (the three powers of a cross-section on three lanes), window test as float compares
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_lanes.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_continue.py tests/test_lowercolorado_lakes.py -m gpu -x -q > gpurun_out/pytest_gpu_lanes.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_lanes.log)" >> $B
ab() { local name=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-trip-order --verify-segments 20000 "$@" > gpurun_out/ab_${name}.json 2> gpurun_out/ab_${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/ab_${name}.json) $(python -c "import json;d=json.loads(open('gpurun_out/ab_${name}.json').read().strip().splitlines()[-1]);print(d['verify']['hash'],d['verify']['mismatches'])" 2>&1 | tail -1)" >> $B
}
ab lanes
ab lanes_off --opt march_lane_groups=0
cat $B
