#!/bin/bash
# Runs each GPU test file in its own process (a trapped kernel poisons the CUDA context of its process only)
# under a timeout, and collects logs under gpurun_out/.  Usage: tests/run_gpu_suite.sh [file ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
FILES="$@"
if [ -z "$FILES" ]; then FILES=$(ls tests/test_gpu_*.py); fi
rc_all=0
for f in $FILES; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -x --timeout=600 > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $name rc=$rc: $(tail -n 1 gpurun_out/$name.log)"
  if [ $rc -ne 0 ]; then rc_all=1; grep -E "^(FAILED|ERROR)|Error|error|assert|pfn:" gpurun_out/$name.log | head -n 25; fi
done
exit $rc_all
