#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=100 -k "transformer_matches_oracle or device_greedy or mid_shape" > gpurun_out/r2j_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2j_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
export LLMF90_PF_LEAD=4
timeout 200 python tools/sweep_env.py tinyllama f32 LLMF90_TILE_WARPS 1 2 3 4 2>&1 | grep -v "^$" | tee gpurun_out/r2j_sweep_f32_g.txt
timeout 200 python tools/sweep_env.py tinyllama f16 MULTI LLMF90_TILE_WARPS=1,LLMF90_SLOT_BYTES=16384 LLMF90_TILE_WARPS=2,LLMF90_SLOT_BYTES=16384 LLMF90_TILE_WARPS=2,LLMF90_SLOT_BYTES=32768 LLMF90_TILE_WARPS=3,LLMF90_SLOT_BYTES=32768 2>&1 | grep -v "^$" | tee gpurun_out/r2j_sweep_f16_g.txt
unset LLMF90_SLOT_BYTES
timeout 300 python tools/sweep_env.py llama2-7b f16 LLMF90_TILE_WARPS 1 2 3 4 2>&1 | grep -v "^$" | tee gpurun_out/r2j_sweep_7bf16_g.txt
timeout 300 python tools/sweep_env.py llama2-7b q4_0 LLMF90_TILE_WARPS 1 2 3 2>&1 | grep -v "^$" | tee gpurun_out/r2j_sweep_7bq4_g.txt
