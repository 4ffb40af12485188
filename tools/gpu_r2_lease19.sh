#!/bin/bash
# Round 2, lease 19: A/B of a library VARIANT (-DTRT_DATAFLOW_FASTDIV: the dataflow kernel's trips with the inline fast-path division)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_fdiv3.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
export TROUTE_B200_LIB=$PWD/t-route_b200/troute_b200/lib/variants/libtroute_b200_fastdiv.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "network_bits or suite or kat or fastpath" > gpurun_out/pytest_gpu_fdiv3.log 2>&1; echo "pytest (variant) rc=$? $(tail -1 gpurun_out/pytest_gpu_fdiv3.log)" >> $B
ab() { local name=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-trip-order --verify-segments 20000 "$@" > gpurun_out/ab_${name}.json 2> gpurun_out/ab_${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/ab_${name}.json) $(python -c "import json;d=json.loads(open('gpurun_out/ab_${name}.json').read().strip().splitlines()[-1]);print(d['verify']['hash'],d['verify']['mismatches'])" 2>&1 | tail -1)" >> $B
}
ab dataflow_fastdiv
unset TROUTE_B200_LIB
ab dataflow_ieee
cat $B
