#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=120 > gpurun_out/r3h_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r3h_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
bash tools/ms_per_token.sh
G=LLMF90_TILE_WARPS
timeout 300 python tools/sweep_env.py tinyllama f16 MULTI $G=3 $G=6 2>&1 | grep ms/token
timeout 300 python tools/sweep_env.py llama2-7b f16 MULTI $G=4 $G=12 2>&1 | grep ms/token
