#!/bin/bash
# tools/ab_lib.sh libA.so libB.so ... -- same-box A/B of builds of libregat.so (REGAT_LIB), device-resident ms/step, 3 interleaved rounds
for round in 1 2 3; do
  for lib in "$@"; do
    ms=$(REGAT_LIB=$PWD/$lib python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4f' % d['ms_per_step'])")
    echo "round $round $lib ms_per_step $ms"
  done
done
