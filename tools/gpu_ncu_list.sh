#!/bin/bash
# tools/gpu_ncu_list.sh <tag> -- ncu launch list (gpu__time_duration) of the default bench command, shares only (cold cache, serialised)
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/${tag}_ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/launch_table.py gpurun_out/${tag}_launches.csv 2>&1 | tail -60
