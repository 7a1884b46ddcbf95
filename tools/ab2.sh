#!/bin/bash
# tools/ab2.sh dirA dirB: 1-GPU and 2-GPU step time of two builds, interleaved
for round in 1 2; do
  for d in "$@"; do
    (cd $d && timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$d 1gpu', round(d['ms_per_step'],4))")
    (cd $d && timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$d 2gpu', round(d['ms_per_step'],4))")
  done
done
