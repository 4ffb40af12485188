#!/bin/bash
# One GPU-box session: first light, GPU tests, bench, ncu launch list + full capture.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
STAGE=${1:-all}
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
