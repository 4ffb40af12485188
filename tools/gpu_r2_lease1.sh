#!/bin/bash
# Round 2, lease 1: Fortran probe, error text of the sharded-nudging test, the A/Bs built at the end of round 1,
# first diffusive bench line + ncu capture.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
{ nproc; free -g | head -2; nvidia-smi -L; } > $B 2>&1
{ echo "== fortran probe"; for c in gfortran flang flang-new nvfortran pgfortran ifort ifx f2c f77 f95 g77 lfortran; do printf "%s: " $c; command -v $c || echo no; done;
  ls /usr/bin/*fortran* /usr/lib/gcc/x86_64-linux-gnu/*/f951 /opt/nvidia/hpc_sdk 2>&1 | head; find / -name "f951" -o -name "libgfortran.so*" 2>/dev/null | head; } >> $B 2>&1
export TRT_TEST_STRICT=1
TRT_TEST_OPEN_ISSUES=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA --tb=long -k sharded_nudging > gpurun_out/pytest_open_issues.log 2>&1; echo "open issues rc=$?" >> $B
ab() { local name=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > "gpurun_out/ab_${name}.json" 2> "gpurun_out/ab_${name}.err"
  echo "ab ${name} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${name}.json")" >> $B; }
ab default
ab warp_resync --opt warp_resync=1
ab trip_totals --trip-buckets 1
ab no_trip_order --no-trip-order
ab march_group1 --opt march_group=1
ab march_group2 --opt march_group=2
make -C t-route_b200/csrc -B EXTRA=-DTRT_DATAFLOW_MIN_BLOCKS=3 > gpurun_out/build_minblocks3.log 2>&1 && ab dataflow_minblocks3
make -C t-route_b200/csrc -B > gpurun_out/build_default.log 2>&1; echo "rebuild default rc=$?" >> $B
timeout 900 python bench.py --workload diffusive --steps 3 --warmup 1 > gpurun_out/bench_diffusive.json 2> gpurun_out/bench_diffusive.err; echo "bench diffusive rc=$?" >> $B
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"time_loop_kernel" -c 1 -f -o gpurun_out/prof_diffusive \
   python bench.py --workload diffusive --domains 148 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_diffusive.log 2>&1; echo "ncu diffusive rc=$?" >> $B
cat $B
