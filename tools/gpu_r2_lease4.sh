#!/bin/bash
# Round 2, lease 4: the solve's read-only state in shared memory (no spills at 64 registers): parity, A/B (calibration on
# another storm / on the timed storm), ncu capture.  Also repeats the Fortran probe and keeps its output.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
name=${1:-r02_smemstate}
{ nproc; free -g | head -2; nvidia-smi -L; } > $B 2>&1
{ echo "== fortran probe on the GPU box ($(date -u +%FT%TZ))"; for c in gfortran flang flang-new nvfortran pgfortran ifort ifx f2c f77 f95 g77 lfortran; do printf "%s: " $c; command -v $c || echo no; done;
  ls /usr/bin/*fortran* /usr/lib/gcc/x86_64-linux-gnu/*/f951 /opt/nvidia/hpc_sdk 2>&1 | head; find / \( -name "f951" -o -name "libgfortran.so*" \) -not -path "/proc/*" 2>/dev/null | head; } > gpurun_out/fortran_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_continue.py -m gpu -x -q > gpurun_out/pytest_${name}.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_${name}.log)" >> $B
ab() { local n=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-verify "$@" > "gpurun_out/ab_${n}.json" 2> "gpurun_out/ab_${n}.err"
  echo "ab ${n} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${n}.json")" >> $B; }
ab ${name}
ab ${name}_samestorm --calibrate-on same-storm
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"dataflow_kernel|march_kernel|finalize_kernel" -s 3 -c 3 -f -o gpurun_out/prof_${name} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify --no-trip-order > gpurun_out/ncu_${name}.log 2>&1; echo "ncu rc=$?" >> $B
cat $B
