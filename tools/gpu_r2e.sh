#!/bin/bash
# round 2, call E: L2 prefetch cursor: quick parity check, then sweeps of the prefetch distance / pace
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "transformer_matches_oracle or device_greedy or mid_shape" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
timeout 300 python tools/sweep_env.py tinyllama f32 LLMF90_PF_LEAD 0 2 5 10 20 40 2>&1 | grep -v "^$" | tee gpurun_out/r2e_sweep_f32_lead.txt
timeout 300 python tools/sweep_env.py tinyllama f32 MULTI LLMF90_PF_LEAD=10,LLMF90_PACE=30 LLMF90_PF_LEAD=10,LLMF90_PACE=34 LLMF90_PF_LEAD=10,LLMF90_PACE=42 LLMF90_PF_LEAD=10,LLMF90_PACE=50 2>&1 | tee gpurun_out/r2e_sweep_f32_pace.txt
unset LLMF90_PACE
timeout 300 python tools/sweep_env.py tinyllama f16 LLMF90_PF_LEAD 0 5 10 20 2>&1 | tee gpurun_out/r2e_sweep_f16_lead.txt
timeout 300 python tools/sweep_env.py llama2-7b q4_0 LLMF90_PF_LEAD 0 8 16 32 2>&1 | tee gpurun_out/r2e_sweep_7bq4_lead.txt
timeout 120 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2e_trace_tinyllama_f32.txt 2>&1; cat gpurun_out/r2e_trace_tinyllama_f32.txt
