#!/bin/bash
S=LLMF90_MAX_SLOTS; G=LLMF90_TILE_WARPS
timeout 500 python tools/sweep_env.py llama2-7b q4_0 MULTI $S=3,$G=6 $S=3,$G=4 $S=3,$G=3 $S=4,$G=6 $S=4,$G=4 $S=3,$G=6,4,6,4,6 $S=3,$G=4,4,6,6,6 2>&1 | grep ms/token
