#!/bin/bash
mkdir -p gpurun_out
M=${1:-tinyllama}; W=${2:-f32}
run() { echo "== $*"; env "$@" timeout 300 python tools/prof_phases.py $M $W 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_token'],4), {k: round(v,3) for k,v in d['phase_ms_per_token'].items()})"; }
run LLMF90_PACE=38
run LLMF90_PACE=0
run LLMF90_PACE=20
run LLMF90_PACE=30
run LLMF90_MAX_SLOTS=4
run LLMF90_MAX_SLOTS=5 LLMF90_PACE=0
run LLMF90_CONS_WARPS=8
run LLMF90_CONS_WARPS=4
