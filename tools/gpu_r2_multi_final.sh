#!/bin/bash
# Round 2, final multi-GPU check of the code as committed.  N = 2: tests/test_gpu_multi.py + the conus line; N = 8: the conus line.
# Run with: gpurun --gpus N -- bash tools/gpu_r2_multi_final.sh N
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
B=gpurun_out/box_multi_final_n$N.txt
{ nproc; nvidia-smi -L | head -8; } > $B 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -rA > gpurun_out/pytest_gpu_multi_final.log 2>&1; echo "pytest multi rc=$? $(tail -1 gpurun_out/pytest_gpu_multi_final.log)" >> $B
fi
run() { local name=$1; shift
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@" > gpurun_out/${name}.json 2> gpurun_out/${name}.err
  echo "${name} rc=$? $(python tools/ab_line.py gpurun_out/${name}.json)" >> $B
  python - >> $B <<PY
import json
try:
    d = json.loads(open("gpurun_out/${name}.json").read().strip().splitlines()[-1])
    v = d.get("verify") or {}
    print("   hash", v.get("hash"), "mismatches", v.get("mismatches"), "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("ms_per_step"), d["config"].get("sharding"))
except Exception as e:
    print("   unreadable", e)
PY
  tail -2 gpurun_out/${name}.err | grep -v "^\*\|OMP_NUM" >> $B
}
run bench_final_n$N --steps 3 --warmup 3 --verify-segments 20000
cat $B
