#!/bin/bash
# One GPU-box session: first light, GPU tests, bench, ncu launch list + full capture.  Everything lands in gpurun_out/.
# Stages: light smoke test open diffusive bench ncu (default: all).  Round 2 starts with: gpu_round.sh "diffusive open"
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
STAGE=${1:-all}
export TRT_TEST_STRICT=1   # first_light tests (tests/conftest.py) fail for real here instead of being reported as xfailed
has() { [[ $STAGE == all || " $STAGE " == *" $1 "* ]]; }
{ nproc; free -g; nvidia-smi -L; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv; } > gpurun_out/box.txt 2>&1
if has light; then
  timeout 900 python tools/gpu_first_light.py > gpurun_out/first_light.log 2>&1; echo "first_light rc=$?" >> gpurun_out/box.txt
fi
if has smoke; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/box.txt
fi
if has test; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/box.txt
fi
if has open; then
  # open issues (skipped by default): full tracebacks, so that the failure of the sharded-nudging test can be read
  TRT_TEST_OPEN_ISSUES=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA --tb=long -k sharded_nudging \
      > gpurun_out/pytest_open_issues.log 2>&1; echo "open issues rc=$?" >> gpurun_out/box.txt
fi
if has diffusive; then
  # first GPU run of the diffusive-wave solver: its own tests with full tracebacks, then its bench lines
  timeout 900 python -m pytest tests/test_zz_gpu_diffusive.py -m gpu -q -rA --tb=long > gpurun_out/pytest_diffusive.log 2>&1
  echo "diffusive tests rc=$?" >> gpurun_out/box.txt
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_zz_gpu_diffusive.py -m gpu -q -x -k "small or uniform" \
      > gpurun_out/sanitizer_diffusive.log 2>&1; echo "diffusive memcheck rc=$?" >> gpurun_out/box.txt
  timeout 900 python bench.py --workload diffusive --steps 3 --warmup 1 > gpurun_out/bench_diffusive.json 2> gpurun_out/bench_diffusive.err
  echo "bench diffusive rc=$?" >> gpurun_out/box.txt
  timeout 600 python bench.py --workload diffusive --impl reference --steps 1 --warmup 0 > gpurun_out/bench_diffusive_ref.json 2>> gpurun_out/bench_diffusive.err
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"time_loop_kernel" -c 1 -f -o gpurun_out/prof_diffusive \
      python bench.py --workload diffusive --domains 148 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_diffusive.log 2>&1
  echo "ncu diffusive rc=$?" >> gpurun_out/box.txt
fi
if has bench; then
  timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/box.txt
  timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> gpurun_out/box.txt
fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/box.txt
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"dataflow_kernel|march_kernel" -s 2 -c 2 -f -o gpurun_out/prof_wavefront \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/box.txt
fi
tail -5 gpurun_out/box.txt
