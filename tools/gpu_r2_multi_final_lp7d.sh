#!/bin/bash
# Round 2, final: config 5 (conus + 5,000 level pools, 7 windows x 288 steps, state handed over on the devices) on N GPUs.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
B=gpurun_out/box_multi_final_lp7d_n$N.txt
{ nproc; nvidia-smi -L | head -8; } > $B 2>&1
name=bench_final_lp7d_n$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload conus-lp7d --steps 2 --warmup 1 --verify-segments 20000 --no-e2e > gpurun_out/${name}.json 2> gpurun_out/${name}.err
echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/${name}.json)" >> $B
python -c "import json; d=json.loads(open('gpurun_out/${name}.json').read().strip().splitlines()[-1]); print('   hash', d['verify']['hash'], 'mismatches', d['verify']['mismatches'], d['config'].get('sharding'))" >> $B 2>&1
cat $B
