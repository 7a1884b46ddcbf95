#!/bin/bash
# tools/round_start.sh -- ONE gpurun call that re-establishes the measured state at the start of a round:
#   GPU test suite, smoke, the default bench line, the launch list of the same command, and an `ncu --set full` capture of the
#   three kernels DESIGN.md section 9 names (dominant GEMM, fused attention forward, attention backward).
# Usage (from the repo root, in the build container):
#   gpurun --timeout 900 -- 'bash tools/round_start.sh r02'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q -p no:cacheprovider > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gpu_tests.log
tail -3 $out/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 30 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"
tail -c 600 $out/${tag}_bench_1gpu.json
# launch list: cold-cache, serialised -- shares only (B200_PROFILING.md)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_list.log 2>&1; echo "ncu list rc=$?"
# full captures of the kernels that bound the step: the three attention kernels, and the first 8 tcgen05 products (qs, uqe, weff,
# v2out, s, Q, KV, pv) of the first step
ncu --set full --clock-control none --import-source on -k "regex:geoattn_fwd_fast_kernel|attn_bwd_bf16_kernel|geo_bwd_kernel" \
    -c 3 -o $out/${tag}_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs > $out/${tag}_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:gemm_tc_kernel" \
    -c 8 -o $out/${tag}_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs > $out/${tag}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
