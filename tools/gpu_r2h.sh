#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=100 -k "transformer_matches_oracle or device_greedy or mid_shape" > gpurun_out/r2h_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2h_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python tools/sweep_env.py tinyllama f32 LLMF90_PF_LEAD 0 4 10 2>&1 | grep -v "^$" | tee gpurun_out/r2h_sweep_f32_lead.txt
unset LLMF90_PF_LEAD
bash tools/ms_per_token.sh
LLMF90_PF_LEAD=4 timeout 150 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2h_trace_tinyllama_f32.txt 2>&1; cat gpurun_out/r2h_trace_tinyllama_f32.txt
