#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout=120 > gpurun_out/r3c_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -6 gpurun_out/r3c_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
bash tools/ms_per_token.sh
timeout 200 python tools/sweep_env.py tinyllama f16 LLMF90_TILE_WARPS 2 3 4 6 2>&1 | grep ms/token
timeout 300 python tools/sweep_env.py llama2-7b f16 LLMF90_TILE_WARPS 2 3 4 6 2>&1 | grep ms/token
timeout 200 python tools/sweep_env.py tinyllama f32 LLMF90_TILE_WARPS 2 3 4 6 2>&1 | grep ms/token
