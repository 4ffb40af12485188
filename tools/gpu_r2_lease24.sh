#!/bin/bash
# Round 2, lease 24: lanes per marching warp -- the occupancy rule (auto) against explicit 1 / 2 lanes, T = 288 and T = 2,016
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_marchgroup.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "marching or network_bits" > gpurun_out/pytest_marchgroup.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_marchgroup.log)" >> $B
ab() { local name=$1; shift
  timeout 900 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-trip-order --verify-segments 20000 "$@" > gpurun_out/ab_${name}.json 2> gpurun_out/ab_${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/ab_${name}.json) $(python -c "import json;d=json.loads(open('gpurun_out/ab_${name}.json').read().strip().splitlines()[-1]);print(d['verify']['hash'],d['verify']['mismatches'])" 2>&1 | tail -1)" >> $B
}
ab mg_auto
ab mg_1 --opt march_group=1
ab mg_2 --opt march_group=2
ab mg_auto_onecall --workload conus-lp7d --windows 1 --nsteps 2016 --steps 2 --warmup 1
ab mg_1_onecall --workload conus-lp7d --windows 1 --nsteps 2016 --steps 2 --warmup 1 --opt march_group=1
cat $B
