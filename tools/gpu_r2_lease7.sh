#!/bin/bash
# Round 2, lease 7: where the time of the end-to-end call goes (TRT_TIMELINE), number of time chunks.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
for c in 4 3 6 8; do
  TRT_TIMELINE=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-trip-order --opt route_chunks=$c > gpurun_out/e2e_chunks$c.json 2> gpurun_out/e2e_chunks$c.err
  echo "chunks $c rc=$? e2e=$(python -c "import json,sys; d=json.loads(open('gpurun_out/e2e_chunks$c.json').read().strip().splitlines()[-1]); print(d['e2e']['ms_per_step'], d['ms_per_step'])")" >> $B
  grep "trt timeline" gpurun_out/e2e_chunks$c.err | tail -1 >> $B
done
cat $B
