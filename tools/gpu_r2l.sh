#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=100 -k "transformer_matches_oracle or device_greedy or mid_shape" > gpurun_out/r2l_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2l_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
export LLMF90_PF_LEAD=4
timeout 200 python tools/sweep_env.py tinyllama f32 LLMF90_TILE_WARPS 3 2,3,3,3,3 2,4,3,3,3 2,4,3,4,3 2,4,4,4,4 2>&1 | grep -v "^$" | tee gpurun_out/r2l_sweep_f32_g.txt
timeout 200 python tools/sweep_env.py tinyllama f16 LLMF90_TILE_WARPS 3 2,3,3,3,3 2,4,3,4,3 2>&1 | grep -v "^$" | tee gpurun_out/r2l_sweep_f16_g.txt
timeout 300 python tools/sweep_env.py llama2-7b q4_0 MULTI LLMF90_TILE_WARPS=4,LLMF90_SLOT_BYTES=36864 LLMF90_TILE_WARPS=4,LLMF90_SLOT_BYTES=18432 LLMF90_TILE_WARPS=6,LLMF90_SLOT_BYTES=36864 2>&1 | grep -v "^$" | tee gpurun_out/r2l_sweep_7bq4_g.txt
unset LLMF90_TILE_WARPS LLMF90_SLOT_BYTES
timeout 150 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2l_trace_tinyllama_f32.txt 2>&1; cat gpurun_out/r2l_trace_tinyllama_f32.txt
