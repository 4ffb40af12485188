#!/bin/bash
# First GPU session of round 2, one gpurun call (~25 min of box time):
#   gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh'
# 1. everything that has never run on a device (tools/gpu_round.sh "diffusive open": diffusive tests with tracebacks,
#    compute-sanitizer, its bench lines and ncu capture; the sharded-nudging test with its error text), then the strict
#    first-light tests (warm restart, time-resolved trip counters);
# 2. the Muskingum-Cunge A/Bs of DESIGN.md "Order of work" (one bench line each, resident number only);
# 3. the default bench + reference arm, and the GPU test suite as the driver runs it.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export TRT_TEST_STRICT=1
bash tools/gpu_round.sh "diffusive open"
timeout 900 python -m pytest tests/test_restart_continuity.py tests/test_route_windows.py tests/test_lowercolorado_lakes.py tests/test_trip_order.py tests/test_gpu_parity.py -m gpu -k 'restart or windows or reservoirs or trip or warp_resync' -q -rA --tb=long \
    > gpurun_out/pytest_first_light.log 2>&1; echo "first-light tests rc=$?" >> gpurun_out/box.txt

# the Muskingum-Cunge parity suite once more with the warp re-synchronisation switched on for every network
TRT_OPTIONS=warp_resync=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q \
    > gpurun_out/pytest_warp_resync.log 2>&1; echo "parity suite under warp_resync rc=$?" >> gpurun_out/box.txt

ab() {   # name, bench arguments...
  local name=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > "gpurun_out/ab_${name}.json" 2> "gpurun_out/ab_${name}.err"
  echo "ab ${name} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${name}.json")" >> gpurun_out/box.txt
}
ab default
ab warp_resync --opt warp_resync=1
ab trip_totals --trip-buckets 1
ab no_trip_order --no-trip-order
ab march_group1 --opt march_group=1
ab march_group2 --opt march_group=2
ab march_group1_deep4096 --opt march_group=1 --deep-lanes 4096
# 80-register dataflow kernel (no spills, 24 warps per SM); rebuilt in place, default build restored afterwards
make -C t-route_b200/csrc -B EXTRA=-DTRT_DATAFLOW_MIN_BLOCKS=3 > gpurun_out/build_minblocks3.log 2>&1 && ab dataflow_minblocks3
make -C t-route_b200/csrc -B > gpurun_out/build_default.log 2>&1; echo "rebuild default rc=$?" >> gpurun_out/box.txt

unset TRT_TEST_STRICT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest (driver style) rc=$?" >> gpurun_out/box.txt
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/box.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> gpurun_out/box.txt
cat gpurun_out/box.txt
