#!/bin/bash
# Round 2, lease 5: summation-order fix of the 5th..nth upstream neighbour, shared trapezoid term, expensive-first trip order.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
name=${1:-r02_lpt}
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${name}.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_${name}.log)" >> $B
ab() { local n=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-verify "$@" > "gpurun_out/ab_${n}.json" 2> "gpurun_out/ab_${n}.err"
  echo "ab ${n} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${n}.json")" >> $B; }
ab ${name}
ab ${name}_samestorm --calibrate-on same-storm
TRT_TRIP_ORDER=asc ab ${name}_samestorm_asc --calibrate-on same-storm
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"dataflow_kernel" -s 4 -c 1 -f -o gpurun_out/prof_${name}_samestorm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify --calibrate-on same-storm > gpurun_out/ncu_${name}.log 2>&1; echo "ncu rc=$?" >> $B
cat $B
