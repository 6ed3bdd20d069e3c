#!/bin/bash
mkdir -p gpurun_out
M=${1:-tinyllama}; W=${2:-f32}
run() { echo "== $*"; env "$@" timeout 100 python tools/prof_phases.py $M $W 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_token'],4), {k: round(v,3) for k,v in d['phase_ms_per_token'].items()})"; }
run LLMF90_PF_STAGES=16 LLMF90_PACE=25
run LLMF90_PF_STAGES=16 LLMF90_PACE=20
run LLMF90_PF_STAGES=16 LLMF90_PACE=12
run LLMF90_PF_STAGES=32 LLMF90_PACE=20
run LLMF90_PF_STAGES=48 LLMF90_PACE=20
run LLMF90_PF_STAGES=0 LLMF90_PACE=20
run LLMF90_PF_STAGES=8 LLMF90_PACE=20
