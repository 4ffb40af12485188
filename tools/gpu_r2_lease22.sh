#!/bin/bash
# Round 2, lease 22: the code as committed -- whole GPU suite, smoke(), the default bench line, config 5, the ncu launch list and
# one --set full capture of the three routing kernels (profiles/r02_v10_final).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_final.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_final2.log)" >> $B
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_final2.log)" >> $B
timeout 1200 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$? $(python tools/ab_line.py gpurun_out/bench_final2.json)" >> $B
python -c "import json; d=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1]); print('   e2e', d['e2e'], 'verify', d['verify']['hash'], d['verify']['mismatches'], 'cpu', d['cpu_baseline'])" >> $B
timeout 1500 python bench.py --workload conus-lp7d --steps 2 --warmup 1 > gpurun_out/bench_final2_lp7d.json 2> gpurun_out/bench_final2_lp7d.err; echo "bench lp7d rc=$? $(python tools/ab_line.py gpurun_out/bench_final2_lp7d.json)" >> $B
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/ncu_launches_final2.log 2>&1; echo "ncu launches rc=$?" >> $B
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"dataflow_kernel|march_kernel|finalize_kernel" -s 12 -c 3 -f -o gpurun_out/prof_final2 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/ncu_final2.log 2>&1; echo "ncu full rc=$?" >> $B
cat $B
