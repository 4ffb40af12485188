#!/bin/bash
# Round 2, lease 9: the device-resident model driver (BMI-shaped windows), the LowerColorado hybrid domain timed on the
# device, does a totals-only trip order carry over to another storm.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_continue.py -m gpu -x -q > gpurun_out/pytest_model.log 2>&1; echo "pytest model rc=$? $(tail -1 gpurun_out/pytest_model.log)" >> $B
timeout 900 python tools/gpu_lc_hybrid_timing.py 288 > gpurun_out/lc_hybrid_timing.json 2> gpurun_out/lc_hybrid_timing.err; echo "lc hybrid rc=$? $(cat gpurun_out/lc_hybrid_timing.json)" >> $B
ab() { local n=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-verify "$@" > "gpurun_out/ab_${n}.json" 2> "gpurun_out/ab_${n}.err"
  echo "ab ${n} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${n}.json") uncal=$(python -c "import json; print('%.4g' % json.loads(open('gpurun_out/ab_${n}.json').read().strip().splitlines()[-1])['value_uncalibrated'])")" >> $B; }
ab r02_totals_other_storm --trip-buckets 1
cat $B
