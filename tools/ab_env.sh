#!/bin/bash
# tools/ab_env.sh VAR v1 v2 ... -- same-box A/B of the default bench (device-resident ms/step only) over values of one environment variable,
# two rounds interleaved so that drift shows up as disagreement between rounds
var=$1; shift
for round in 1 2; do
  for v in "$@"; do
    ms=$(env $var=$v python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4f' % d['ms_per_step'])")
    echo "round $round $var=$v ms_per_step $ms"
  done
done
