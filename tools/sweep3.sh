#!/bin/bash
run() { M=$1; W=$2; shift 2; echo "== $M $W $*"; env "$@" timeout 100 python tools/prof_phases.py $M $W --noprof 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_token'],4))"; }
run tinyllama f32 LLMF90_SMEM_PAD=0
run tinyllama f32 LLMF90_SMEM_PAD=24576
run tinyllama f32 LLMF90_SMEM_PAD=49152
run tinyllama f32 LLMF90_MAX_SLOTS=4 LLMF90_SMEM_PAD=0
run tinyllama f32 LLMF90_MAX_SLOTS=4 LLMF90_SMEM_PAD=73728
run tinyllama f32 LLMF90_PACE=32
run tinyllama f32 LLMF90_PACE=44
