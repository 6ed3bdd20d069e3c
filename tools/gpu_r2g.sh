#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=120 > gpurun_out/r2g_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r2g_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
for m in "tinyllama f32 10 64" "llama2-7b q4_0 10 64"; do
  set -- $m
  timeout 150 python tools/prof_trace.py $1 $2 $3 $4 > gpurun_out/r2g_trace_$1_$2_$4.txt 2>&1; cat gpurun_out/r2g_trace_$1_$2_$4.txt | grep -v "^warp 0, first"
done
bash tools/ms_per_token.sh
