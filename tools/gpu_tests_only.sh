#!/bin/bash
# tools/gpu_tests_only.sh <tag> [pytest args] -- GPU tests only, log under gpurun_out/<tag>_gpu_tests.log
tag=${1:-rXX}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s "$@" > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_gpu_tests.log
grep -n "bf16 grads\|bf16 train\|\[dp\]\|argmax\]" gpurun_out/${tag}_gpu_tests.log | cut -c1-400
grep -n "^FAILED\|^ERROR\|passed\|failed" gpurun_out/${tag}_gpu_tests.log | tail -40
