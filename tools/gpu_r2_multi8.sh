#!/bin/bash
# Round 2: 8 GPUs -- the conus line (BASELINE configs[2]) and config 5 (conus + 5,000 level pools, 7 windows of 288 steps with
# the state handed over on the devices), both with their verify object (same result hash as N = 1).
# Run with: gpurun --gpus 8 -- bash tools/gpu_r2_multi8.sh
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_multi8.txt
{ nproc; nvidia-smi -L; nvidia-smi topo -m | head -12; } > $B 2>&1
N=${1:-8}
run() { local name=$1; shift
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N "$@" > gpurun_out/${name}.json 2> gpurun_out/${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/${name}.json)" >> $B
  python - >> $B <<PY
import json
try:
    d = json.loads(open("gpurun_out/${name}.json").read().strip().splitlines()[-1])
    v = d.get("verify") or {}
    print("   hash", v.get("hash"), "mismatches", v.get("mismatches"), "e2e", (d.get("e2e") or {}).get("value"), "uncal", d.get("value_uncalibrated"), d["config"].get("sharding"), d["config"].get("host_placement"))
except Exception as e:
    print("   unreadable", e)
PY
}
run bench_r02_n$N --steps 3 --warmup 3
run bench_r02_lp7d_n$N --workload conus-lp7d --steps 2 --warmup 1 --verify-segments 20000
cat $B
