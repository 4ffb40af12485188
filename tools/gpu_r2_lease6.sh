#!/bin/bash
# Round 2, lease 6: where the dataflow kernel's time goes stage by stage (narrow tail), split level A/B.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
{ nproc; nvidia-smi -L; } > $B 2>&1
timeout 600 python tools/gpu_stage_profile.py 4 > gpurun_out/stage_profile_mode4_r02.log 2>&1; echo "stage profile rc=$?" >> $B
ab() { local n=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-verify "$@" > "gpurun_out/ab_${n}.json" 2> "gpurun_out/ab_${n}.err"
  echo "ab ${n} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${n}.json")" >> $B; }
ab r02_deep16k --deep-lanes 16384
ab r02_deep32k --deep-lanes 32768
ab r02_deep64k --deep-lanes 65536
cat $B; cat gpurun_out/stage_profile_mode4_r02.log
