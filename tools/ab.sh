#!/bin/bash
# A/B timing of several builds of the repo inside one GPU session: tools/ab.sh dirA dirB ... (each prints ms/step, 2 rounds)
for round in 1 2; do
  for d in "$@"; do
    (cd $d && python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$d', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))")
  done
done
