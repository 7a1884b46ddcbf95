#!/bin/bash
# tools/gpu_ncu_full.sh <tag> <kernel regex> [count] -- one `ncu --set full` capture of the named kernels inside the default bench command
tag=${1:-rXX}; rx=${2:-geoattn_fwd_fast_kernel}; cnt=${3:-3}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip 6 -c $cnt -o gpurun_out/${tag}_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${tag}_full.ncu-rep
