#!/bin/bash
# Round 2, lease 11 (1 GPU): the small config-5 case of the 2-GPU lease on one GPU (same result hash expected), and config 5 as
# one call with the marching lanes-per-warp rule that counts live segments.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
ab() { local n=$1; shift
  timeout 1200 python bench.py --no-cpu-baseline --no-e2e "$@" > "gpurun_out/ab_${n}.json" 2> "gpurun_out/ab_${n}.err"
  echo "ab ${n} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${n}.json") hash=$(python -c "import json; d=json.loads(open('gpurun_out/ab_${n}.json').read().strip().splitlines()[-1]); v=d.get('verify') or {}; print(v.get('hash'), 'mismatches', v.get('mismatches'))")" >> $B; tail -2 "gpurun_out/ab_${n}.err" >> $B; }
ab r02_n1_lp7d_small --workload conus-lp7d --segments 400000 --windows 3 --nsteps 96 --steps 2 --warmup 1 --verify-segments 8000 --no-trip-order
ab r02_lp7d_onecall_g --workload conus-lp7d --windows 1 --nsteps 2016 --steps 2 --warmup 1 --no-verify --no-trip-order
cat $B
