#!/bin/bash
# Round 2, lease 17: marching lanes with the inline fast-path division (McDivFast) -- whole GPU suite, then the bench line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_fdiv.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_fdiv.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_fdiv.log)" >> $B
ab() { local name=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-trip-order --verify-segments 20000 "$@" > gpurun_out/ab_${name}.json 2> gpurun_out/ab_${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/ab_${name}.json) $(python -c "import json;d=json.loads(open('gpurun_out/ab_${name}.json').read().strip().splitlines()[-1]);print(d['verify']['hash'],d['verify']['mismatches'])" 2>&1 | tail -1)" >> $B
}
ab fdiv
ab fdiv_lp7d --workload conus-lp7d --windows 2
cat $B
