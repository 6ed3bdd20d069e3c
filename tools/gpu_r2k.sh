#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=100 -k "transformer_matches_oracle or device_greedy or mid_shape" > gpurun_out/r2k_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2k_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 200 python tools/sweep_env.py tinyllama f16 MULTI LLMF90_TILE_WARPS=3,LLMF90_PF_LEAD=4 LLMF90_TILE_WARPS=4,LLMF90_PF_LEAD=4 LLMF90_TILE_WARPS=3,LLMF90_PF_LEAD=8 LLMF90_TILE_WARPS=3,LLMF90_PF_LEAD=16 2>&1 | grep -v "^$" | tee gpurun_out/r2k_sweep_f16.txt
unset LLMF90_TILE_WARPS
timeout 300 python tools/sweep_env.py llama2-7b f16 LLMF90_PF_LEAD 0 4 8 16 2>&1 | grep -v "^$" | tee gpurun_out/r2k_sweep_7bf16_lead.txt
timeout 300 python tools/sweep_env.py llama2-7b q4_0 MULTI LLMF90_TILE_WARPS=3,LLMF90_PF_LEAD=4 LLMF90_TILE_WARPS=4,LLMF90_PF_LEAD=4 LLMF90_TILE_WARPS=3,LLMF90_PF_LEAD=8,LLMF90_SLOT_BYTES=36864 LLMF90_TILE_WARPS=4,LLMF90_PF_LEAD=8,LLMF90_SLOT_BYTES=36864  2>&1 | grep -v "^$" | tee gpurun_out/r2k_sweep_7bq4.txt
unset LLMF90_TILE_WARPS LLMF90_SLOT_BYTES
LLMF90_PF_LEAD=4 timeout 150 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2k_trace_tinyllama_f32.txt 2>&1; cat gpurun_out/r2k_trace_tinyllama_f32.txt
