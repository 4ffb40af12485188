#!/bin/bash
# Round 2, lease 8: the bench lines of record (1 GPU): conus (configs[2]) with verify / e2e / cpu_baseline, the reference arm,
# conus-lp7d (configs[4]) on one GPU, ncu launch list + one --set full capture of the default configuration.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 1200 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$? $(python tools/ab_line.py gpurun_out/bench_r02.json)" >> $B
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; echo "bench ref rc=$? $(tail -c 600 gpurun_out/bench_r02_ref.json | head -c 300)" >> $B
timeout 1500 python bench.py --workload conus-lp7d --steps 2 --warmup 1 > gpurun_out/bench_r02_lp7d.json 2> gpurun_out/bench_r02_lp7d.err; echo "bench lp7d rc=$? $(python tools/ab_line.py gpurun_out/bench_r02_lp7d.json)" >> $B
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/ncu_launches_r02.log 2>&1; echo "ncu launches rc=$?" >> $B
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"dataflow_kernel|march_kernel|finalize_kernel" -s 12 -c 3 -f -o gpurun_out/prof_r02_final \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/ncu_r02_final.log 2>&1; echo "ncu full rc=$?" >> $B
cat $B
