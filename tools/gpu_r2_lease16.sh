#!/bin/bash
# Round 2, lease 16: dataflow_park_kernel after the footprint work -- parity, A/Bs, a light ncu read-out per variant.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_park2.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 1200 python -m pytest tests/test_gpu_park.py -x -q > gpurun_out/pytest_park.log 2>&1; echo "pytest park rc=$? $(tail -1 gpurun_out/pytest_park.log)" >> $B
ab() { local name=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-trip-order --verify-segments 20000 "$@" > gpurun_out/ab_${name}.json 2> gpurun_out/ab_${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/ab_${name}.json) $(python -c "import json;d=json.loads(open('gpurun_out/ab_${name}.json').read().strip().splitlines()[-1]);print(d['verify']['hash'],d['verify']['mismatches'])" 2>&1 | tail -1)" >> $B
}
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,gpu__time_duration.sum
light() { local name=$1; shift
  timeout 600 ncu --metrics $M --clock-control none -k regex:"dataflow" -s 1 -c 1 --csv --log-file gpurun_out/ncu_light_${name}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify --no-trip-order "$@" > gpurun_out/ncu_light_${name}.log 2>&1
  echo "light ${name} rc=$? $(python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_light_${name}.csv")) if len(r)>10]
print(" ".join(f"{r[-3].split('__')[-1][:34]}={r[-1]}" for r in rows[1:]))
PY
)" >> $B
}
if grep -q "passed" gpurun_out/pytest_park.log && ! grep -q "failed" gpurun_out/pytest_park.log; then
  ab park_off --opt park_max=0 --opt early_max_tiles=0
  ab park_default
  ab park_only --opt early_max_tiles=0
  ab early_only --opt park_max=0
  ab plain_new --opt park_max=0 --opt park_min_tiles=0 --opt early_max_tiles=1
  light park_default
  light plain_new --opt park_max=0 --opt park_min_tiles=0 --opt early_max_tiles=1
fi
cat $B
