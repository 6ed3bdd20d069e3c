#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=120 > gpurun_out/r2o_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r2o_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 150 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2o_trace_tinyllama_f32.txt 2>&1; grep -v "^warp 0, first" gpurun_out/r2o_trace_tinyllama_f32.txt
timeout 150 python tools/prof_trace.py llama2-7b q4_0 10 64 > gpurun_out/r2o_trace_7b_q4.txt 2>&1; tail -3 gpurun_out/r2o_trace_7b_q4.txt
bash tools/ms_per_token.sh
