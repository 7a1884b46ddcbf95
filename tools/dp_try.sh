#!/bin/bash
# tools/dp_try.sh N "ENV1=.. ENV2=.." ...   -- one short N-GPU bench per environment setting, each under its own timeout
N=$1; shift
port=29600
for envs in "$@"; do
  port=$((port+1))
  echo "== $envs"
  env $envs timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))" || echo failed
done
