#!/bin/bash
run() { M=$1; W=$2; shift 2; echo "== $M $W $*"; env "$@" timeout 100 python tools/prof_phases.py $M $W --noprof 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_token'],4))"; }
for S in 4 5 6; do run tinyllama f32 LLMF90_MAX_SLOTS=$S; done
run tinyllama f32 LLMF90_PACE=32
run tinyllama f32 LLMF90_LL_REP=2
run tinyllama f32 LLMF90_LL_REP=8
for S in 4 6; do run llama2-7b f16 LLMF90_MAX_SLOTS=$S; run llama2-7b q4_0 LLMF90_MAX_SLOTS=$S; run tinyllama f16 LLMF90_MAX_SLOTS=$S; done
