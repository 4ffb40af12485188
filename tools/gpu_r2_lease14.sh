#!/bin/bash
# Round 2, lease 14: dataflow_park_kernel -- parity of every park / early setting, then A/Bs on the bench workload.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_park.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 1200 python -m pytest tests/test_gpu_park.py -x -q > gpurun_out/pytest_park.log 2>&1; echo "pytest park rc=$? $(tail -1 gpurun_out/pytest_park.log)" >> $B
ab() { local name=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-trip-order --verify-segments 20000 "$@" > gpurun_out/ab_${name}.json 2> gpurun_out/ab_${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/ab_${name}.json) $(python -c "import json;d=json.loads(open('gpurun_out/ab_${name}.json').read().strip().splitlines()[-1]);print(d['verify']['hash'],d['verify']['mismatches'])" 2>&1 | tail -1)" >> $B
}
if grep -q "passed" gpurun_out/pytest_park.log && ! grep -q "failed" gpurun_out/pytest_park.log; then
  ab park_off --opt park_max=0 --opt early_max_tiles=0
  ab park_default
  ab park_only --opt early_max_tiles=0
  ab early_only --opt park_max=0
  ab park4 --opt park_max=4
  ab park12 --opt park_max=12
fi
cat $B
