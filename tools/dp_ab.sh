#!/bin/bash
# 2-GPU A/B of env switches: device-resident ms/step
run() {
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus ${NG:-2} --steps 40 --warmup 5 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('%.4f' % d['ms_per_step'])"
}
for cfg in "X=1" "REGAT_TC_PDL=0" "REGAT_OPT_PDL=0" "REGAT_TC_PDL=0 REGAT_OPT_PDL=0" "REGAT_OPT_PRIO=low" "X=1"; do
  echo "$cfg -> $(run $cfg)"
done
