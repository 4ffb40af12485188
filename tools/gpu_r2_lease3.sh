#!/bin/bash
# Round 2: first device run of the tiled flow state + TMA-staged tile records: sanitizer, GPU suite, smoke, bench, ncu.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $B
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $B
tail -5 gpurun_out/pytest_gpu.log >> $B
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $B
ab() { local name=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > "gpurun_out/ab_${name}.json" 2> "gpurun_out/ab_${name}.err"
  echo "ab ${name} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${name}.json")" >> $B; }
ab tiled_default
ab tiled_no_trip_order --no-trip-order
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"dataflow_kernel|march_kernel|finalize_kernel" -s 3 -c 3 -f -o gpurun_out/prof_r02_tiled \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $B
cat $B
