#!/bin/bash
# Round 2, lease 13: the whole GPU suite, smoke() and the default bench line on the code as committed.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_final.log)" >> $B
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_final.log)" >> $B
timeout 1200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$? $(python tools/ab_line.py gpurun_out/bench_final.json)" >> $B
python -c "import json; d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print('   e2e', d['e2e'], 'verify', d['verify']['hash'], d['verify']['mismatches'])" >> $B
cat $B
