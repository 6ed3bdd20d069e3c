#!/bin/bash
mkdir -p gpurun_out
export LLMF90_PF_LEAD=4
timeout 200 python tools/sweep_env.py tinyllama f32 LLMF90_SLOT_BYTES 16384 24576 32768 2>&1 | grep -v "^$" | tee gpurun_out/r2i_sweep_f32_slot.txt
timeout 200 python tools/sweep_env.py tinyllama f16 LLMF90_SLOT_BYTES 8192 16384 32768 2>&1 | grep -v "^$" | tee gpurun_out/r2i_sweep_f16_slot.txt
timeout 300 python tools/sweep_env.py llama2-7b f16 LLMF90_SLOT_BYTES 16384 32768 2>&1 | grep -v "^$" | tee gpurun_out/r2i_sweep_7bf16_slot.txt
timeout 300 python tools/sweep_env.py llama2-7b q4_0 LLMF90_SLOT_BYTES 9216 18432 36864 2>&1 | grep -v "^$" | tee gpurun_out/r2i_sweep_7bq4_slot.txt
