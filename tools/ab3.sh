#!/bin/bash
# tools/ab3.sh "ENV=.. dir" ...: short 1-GPU bench per (environment, build) pair, two rounds
for round in 1 2; do
  for spec in "$@"; do
    d=${spec##* }; envs=${spec% *}
    (cd $d && env $envs timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$spec', round(d['ms_per_step'],4))")
  done
done
