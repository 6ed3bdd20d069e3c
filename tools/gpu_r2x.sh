#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=120 > gpurun_out/r2x_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r2x_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 200 python tools/sweep_env.py tinyllama f16 MULTI LLMF90_TILE_WARPS=3 LLMF90_TILE_WARPS=2 LLMF90_TILE_WARPS=4 LLMF90_TILE_WARPS=3,LLMF90_SLOT_BYTES=32768 LLMF90_TILE_WARPS=4,LLMF90_SLOT_BYTES=32768 LLMF90_TILE_WARPS=3,LLMF90_MAX_SLOTS=8 2>&1 | grep ms/token
unset LLMF90_TILE_WARPS LLMF90_SLOT_BYTES LLMF90_MAX_SLOTS
timeout 300 python tools/sweep_env.py llama2-7b f16 MULTI LLMF90_TILE_WARPS=3 LLMF90_TILE_WARPS=4 LLMF90_TILE_WARPS=3,LLMF90_MAX_SLOTS=3 LLMF90_TILE_WARPS=4,LLMF90_MAX_SLOTS=3 2>&1 | grep ms/token
