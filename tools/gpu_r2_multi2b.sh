#!/bin/bash
# Round 2: 2 GPUs -- the extended 2-GPU parity test (sharded nudging, sharded multi-window hand-off vs the oracle on every row,
# with and without the overlapped marching kernel) and the overlap A/B at N = 2.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_multi2b.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -rA > gpurun_out/pytest_gpu_multi_n2b.log 2>&1; echo "pytest multi rc=$? $(tail -1 gpurun_out/pytest_gpu_multi_n2b.log)" >> $B
run() { local name=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 "$@" > gpurun_out/${name}.json 2> gpurun_out/${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/${name}.json) hash=$(python -c "import json; d=json.loads(open('gpurun_out/${name}.json').read().strip().splitlines()[-1]); v=d.get('verify') or {}; print(v.get('hash'), 'mismatches', v.get('mismatches'))")" >> $B; tail -2 gpurun_out/${name}.err >> $B; }
run bench_r02_n2_overlap --steps 3 --warmup 3 --no-e2e --opt overlap_march=1
run bench_r02_n2_lp7d_small --workload conus-lp7d --segments 400000 --windows 3 --nsteps 96 --steps 2 --warmup 1 --no-e2e --verify-segments 8000
cat $B
