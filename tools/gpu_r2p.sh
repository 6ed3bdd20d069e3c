#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=100 -k "transformer_matches_oracle or device_greedy or mid_shape" > gpurun_out/r2p_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2p_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
bash tools/ms_per_token.sh
timeout 150 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2p_trace_tinyllama_f32.txt 2>&1; grep -v "^warp 0, first" gpurun_out/r2p_trace_tinyllama_f32.txt
