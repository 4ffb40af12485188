#!/bin/bash
# quick A/B lease: parity suite (MC part) + one or more bench lines.  usage: gpu_ab.sh name [bench args...]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=gpurun_out/box.txt
name=$1; shift
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_continue.py -m gpu -x -q > gpurun_out/pytest_${name}.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_${name}.log)" > $B
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > "gpurun_out/ab_${name}.json" 2> "gpurun_out/ab_${name}.err"
echo "ab ${name} rc=$? $(python tools/ab_line.py "gpurun_out/ab_${name}.json")" >> $B
if [ "${NCU:-0}" = "1" ]; then
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"dataflow_kernel|march_kernel|finalize_kernel" -s 3 -c 3 -f -o gpurun_out/prof_${name} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_${name}.log 2>&1; echo "ncu rc=$?" >> $B
fi
cat $B
