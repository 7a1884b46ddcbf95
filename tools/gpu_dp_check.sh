#!/bin/bash
# tools/gpu_dp_check.sh <tag> <ngpu> -- multi-GPU: the data-parallel equivalence tests (every exchange backend) and the bench line at N GPUs
tag=${1:-rXX}; n=${2:-2}
mkdir -p gpurun_out
if [ -z "${SKIP_TESTS:-}" ]; then
REGAT_TEST_WORLD=$n timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -p no:cacheprovider -s > gpurun_out/${tag}_dp_tests_${n}gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_dp_tests_${n}gpu.log
grep -n "\[dp\]\|passed\|failed\|^FAILED\|^E  " gpurun_out/${tag}_dp_tests_${n}gpu.log | cut -c1-300 | tail -30
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 30 --warmup 5 \
   > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/${tag}_bench_${n}gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench_${n}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","launches_per_step","final_loss","n_gpus")})
    print("dp_check",d.get("dp_check")); print("e2e",d["e2e"]); print("sustained",d.get("sustained")); print("e2e16",d.get("e2e_host_bf16"))
    print("workloads",json.dumps(d["workloads"])[:1800]); print("cfg",d["config"])
except Exception as ex: print("parse failed",ex)
PY
