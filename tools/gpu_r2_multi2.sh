#!/bin/bash
# Round 2: the 2-GPU parity test on two real devices (CUDA-IPC peer stores over NVLink; incl. sharded nudging) and the
# 2-GPU bench line with its verify object (same result hash as N = 1).  Run with: gpurun --gpus 2 -- bash tools/gpu_r2_multi2.sh
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_multi2.txt
{ nproc; nvidia-smi -L; nvidia-smi topo -m | head -6; } > $B 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -rA > gpurun_out/pytest_gpu_multi_n2.log 2>&1; echo "pytest multi rc=$? $(tail -1 gpurun_out/pytest_gpu_multi_n2.log)" >> $B
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench n2 rc=$? $(python tools/ab_line.py gpurun_out/bench_r02_n2.json)" >> $B
python - >> $B <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_n2.json").read().strip().splitlines()[-1])
    print("n2 verify:", d.get("verify"), "e2e:", d.get("e2e"))
except Exception as e:
    print("unreadable", e)
PY
cat $B
