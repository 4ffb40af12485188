#!/bin/bash
# Round 2, lease 27 (last GPU minutes): marching kernel with one CTA per SM (-DTRT_MARCH_MIN_BLOCKS=1, library variant) against two
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_march1.txt
nvidia-smi -L > $B 2>&1
ab() { local name=$1; shift
  timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-trip-order --verify-segments 20000 "$@" > gpurun_out/ab_${name}.json 2> gpurun_out/ab_${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/ab_${name}.json) $(python -c "import json;d=json.loads(open('gpurun_out/ab_${name}.json').read().strip().splitlines()[-1]);print(d['verify']['hash'],d['verify']['mismatches'])" 2>&1 | tail -1)" >> $B
}
TROUTE_B200_LIB=$PWD/t-route_b200/troute_b200/lib/variants/libtroute_b200_march1.so ab march_1cta
cat $B
