#!/bin/bash
P=LLMF90_PACE; F=LLMF90_PF_LEAD; R=LLMF90_LL_REP
timeout 500 python tools/sweep_env.py tinyllama f32 MULTI $P=38 $P=0 $P=30 $P=46 $F=3 $F=6 $F=14 $R=2 $R=8 $P=38 2>&1 | grep ms/token
timeout 300 python tools/sweep_env.py tinyllama f16 MULTI $P=38 $P=30 $P=46 $F=4 $R=8 2>&1 | grep ms/token
timeout 300 python tools/sweep_env.py llama2-7b q4_0 MULTI $P=38 $P=30 $F=4 $R=8 2>&1 | grep ms/token
