#!/bin/bash
# Round 2: 8 GPUs after the reservoir-above-a-cut-edge fix -- config 5 (7 windows of 288 steps, state handed over on the
# devices; the result hash must equal the 1-GPU run's) and the same week as ONE call of 2,016 steps.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_multi8b.txt
{ nproc; nvidia-smi -L | head -8; } > $B 2>&1
N=${1:-8}
run() { local name=$1; shift
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus $N "$@" > gpurun_out/${name}.json 2> gpurun_out/${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/${name}.json)" >> $B
  python - >> $B <<PY
import json
try:
    d = json.loads(open("gpurun_out/${name}.json").read().strip().splitlines()[-1])
    v = d.get("verify") or {}
    print("   hash", v.get("hash"), "mismatches", v.get("mismatches"), "e2e", (d.get("e2e") or {}).get("value"), d["config"].get("sharding"))
except Exception as e:
    print("   unreadable", e)
PY
  tail -2 gpurun_out/${name}.err | grep -v "^\*\|OMP_NUM" >> $B
}
run bench_r02_lp7d_n${N}_fixed --workload conus-lp7d --steps 2 --warmup 1 --verify-segments 20000
run bench_r02_lp7d_onecall_n$N --workload conus-lp7d --windows 1 --nsteps 2016 --steps 2 --warmup 1 --no-e2e --verify-segments 20000
cat $B
