#!/bin/bash
# Round 2, lease 23: is the e2e of lease 22 (235.7 ms against 217 ms before) the box or the code?  PCIe probe + three e2e lines.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box_e2e.txt
{ nproc; nvidia-smi -L; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv; } > $B 2>&1
timeout 300 python tools/gpu_pcie_probe2.py >> $B 2>&1
e2e() { local name=$1; shift
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-trip-order --verify-segments 20000 "$@" > gpurun_out/e2e_${name}.json 2> gpurun_out/e2e_${name}.err
  echo "${name} rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/e2e_${name}.json').read().strip().splitlines()[-1]);print('ms_per_step',round(d['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2))" 2>&1 | tail -1)" >> $B
}
e2e auto
e2e chunks6 --opt route_chunks=6
e2e chunks4 --opt route_chunks=4
TRT_TIMELINE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-trip-order --no-verify > gpurun_out/e2e_timeline.json 2> gpurun_out/e2e_timeline.err
grep -i "chunk" gpurun_out/e2e_timeline.err | tail -14 >> $B
cat $B
