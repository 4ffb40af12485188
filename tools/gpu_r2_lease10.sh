#!/bin/bash
# Round 2, lease 10: marching kernel beside the dataflow kernel (overlap_march): parity, 1-GPU A/B; config 5 as ONE call of
# 2,016 steps (the marching chain is paid once: 2,703 + 2,016 links instead of 7 x (2,703 + 288)).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_overlap.log 2>&1; echo "pytest parity rc=$? $(tail -1 gpurun_out/pytest_overlap.log)" >> $B
ab() { local n=$1; shift
  timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > "gpurun_out/ab_${n}.json" 2> "gpurun_out/ab_${n}.err"
  echo "ab ${n} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${n}.json") hash=$(python -c "import json; d=json.loads(open('gpurun_out/ab_${n}.json').read().strip().splitlines()[-1]); v=d.get('verify') or {}; print(v.get('hash'), 'mismatches', v.get('mismatches'))")" >> $B; tail -3 "gpurun_out/ab_${n}.err" >> $B; }
ab r02_overlap_n1 --opt overlap_march=1 --no-trip-order
ab r02_lp7d_onecall --workload conus-lp7d --windows 1 --nsteps 2016 --steps 2 --warmup 1 --verify-segments 20000 --no-trip-order
ab r02_lp7d_onecall_overlap --workload conus-lp7d --windows 1 --nsteps 2016 --steps 2 --warmup 1 --no-verify --no-trip-order --opt overlap_march=1
cat $B
