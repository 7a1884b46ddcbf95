#!/bin/bash
# tools/gpu_check.sh <tag> -- one gpurun call: GPU test suite, smoke, default bench line (all logs under gpurun_out/<tag>_*)
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -s ${PYTEST_ARGS:-} > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_gpu_tests.log
tail -15 $out/${tag}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 $out/${tag}_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"
tail -c 1500 $out/${tag}_bench_1gpu.err
python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_bench_1gpu.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","launches_per_step","final_loss")})
    print("e2e",d["e2e"]); print("sustained",d.get("sustained")); print("e2e16",d.get("e2e_host_bf16"))
    r=d["roofline"]; print({k:r[k] for k in r if k!="gemm_class"})
    print("attn",d["roofline_attention"]); print("workloads",json.dumps(d["workloads"])[:1500]); print("clocks",d["clocks"])
    print("cpu",d["cpu_baseline"])
    for rec in (r.get("gemm_class") or {}).items(): print(rec)
except Exception as ex: print("parse failed",ex)
PY
