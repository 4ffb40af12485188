#!/bin/bash
# Round 2, lease 25: lanes per marching warp chosen per level; with that, how many levels should march?
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_marchlevels.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "marching or network_bits or chunk" > gpurun_out/pytest_marchlevels.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_marchlevels.log)" >> $B
ab() { local name=$1; shift
  timeout 900 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-trip-order --verify-segments 20000 "$@" > gpurun_out/ab_${name}.json 2> gpurun_out/ab_${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/ab_${name}.json) $(python -c "import json;d=json.loads(open('gpurun_out/ab_${name}.json').read().strip().splitlines()[-1]);print(d['verify']['hash'],d['verify']['mismatches'],d['roofline'].get('first_marching_level'))" 2>&1 | tail -1)" >> $B
}
ab perlevel_8k
ab perlevel_12k --deep-lanes 12288
ab perlevel_16k --deep-lanes 16384
ab perlevel_24k --deep-lanes 24576
ab perlevel_40k --deep-lanes 40960
cat $B
