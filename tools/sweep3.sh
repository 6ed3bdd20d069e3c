#!/bin/bash
run() { M=$1; W=$2; shift 2; echo "== $M $W $*"; env "$@" timeout 100 python tools/prof_phases.py $M $W --noprof 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_token'],4))"; }
for A in 0 64 148; do run llama2-7b q4_0 LLMF90_ATT_ITEMS=$A; run tinyllama f32 LLMF90_ATT_ITEMS=$A; run llama2-7b f16 LLMF90_ATT_ITEMS=$A; done
