#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=120 > gpurun_out/r3i_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r3i_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
bash tools/ms_per_token.sh
for c in "tinyllama f32" "llama2-7b q4_0"; do
  set -- $c
  echo "=== trace $1 $2"
  timeout 200 python tools/prof_trace.py $1 $2 10 64 2>&1 | tail -18 | cut -c1-330
done
