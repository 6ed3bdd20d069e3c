#!/bin/bash
P=LLMF90_PACE
timeout 300 python tools/sweep_env.py tinyllama f32 $P 38 42 46 50 54 38 46 2>&1 | grep ms/token
timeout 200 python tools/sweep_env.py tinyllama f16 $P 38 46 54 2>&1 | grep ms/token
timeout 300 python tools/sweep_env.py llama2-7b f16 $P 38 46 54 2>&1 | grep ms/token
timeout 300 python tools/sweep_env.py llama2-7b q4_0 $P 38 46 54 2>&1 | grep ms/token
