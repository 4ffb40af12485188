#!/bin/bash
# Round 2, lease 15: ncu --set full of dataflow_park_kernel (plain path: no parking, no early publication) next to dataflow_kernel
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_park_ncu.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
cap() { local name=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dataflow" -s 1 -c 1 -f -o gpurun_out/prof_${name} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify --no-trip-order "$@" > gpurun_out/ncu_${name}.log 2>&1; echo "ncu ${name} rc=$?" >> $B
}
cap park_plain --opt park_max=0 --opt park_min_tiles=0 --opt early_max_tiles=1
cap park_on
cap park_old --opt park_max=0 --opt early_max_tiles=0
cat $B
