#!/bin/bash
# Round 2, lease 26: the code as committed at the end of the round -- whole GPU suite, smoke(), the default bench line, config 5.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_final3.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final3.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_final3.log)" >> $B
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final3.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_final3.log)" >> $B
timeout 1200 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo "bench rc=$? $(python tools/ab_line.py gpurun_out/bench_final3.json)" >> $B
python -c "import json; d=json.loads(open('gpurun_out/bench_final3.json').read().strip().splitlines()[-1]); print('   e2e', d['e2e'], 'verify', d['verify']['hash'], d['verify']['mismatches'], 'uncal', d.get('value_uncalibrated'), 'incl_h2d', d.get('value_incl_h2d'), 'roofline', d['roofline']['frac'])" >> $B
timeout 1500 python bench.py --workload conus-lp7d --steps 2 --warmup 1 > gpurun_out/bench_final3_lp7d.json 2> gpurun_out/bench_final3_lp7d.err; echo "bench lp7d rc=$? $(python tools/ab_line.py gpurun_out/bench_final3_lp7d.json)" >> $B
python -c "import json; d=json.loads(open('gpurun_out/bench_final3_lp7d.json').read().strip().splitlines()[-1]); print('   e2e', d['e2e']['value'], 'verify', d['verify']['hash'], d['verify']['mismatches'])" >> $B
cat $B
