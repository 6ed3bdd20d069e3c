#!/bin/bash
# sweep of the producer knobs (L2 prefetch distance, pacing, slot size) on one config
mkdir -p gpurun_out
M=${1:-tinyllama}; W=${2:-f32}
run() { echo "== $*"; env "$@" timeout 300 python tools/prof_phases.py $M $W 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_token'],4), {k: round(v,3) for k,v in d['phase_ms_per_token'].items()})"; }
run LLMF90_PF_STAGES=0
run LLMF90_PF_STAGES=0 LLMF90_PACE=0
run LLMF90_PF_STAGES=0 LLMF90_PACE=32
run LLMF90_PF_STAGES=0 LLMF90_PACE=44
run LLMF90_PF_STAGES=8
run LLMF90_PF_STAGES=16
run LLMF90_PF_STAGES=32
run LLMF90_PF_STAGES=64
run LLMF90_PF_STAGES=32 LLMF90_PACE=44
run LLMF90_PF_STAGES=0 LLMF90_SLOT_BYTES=8192 LLMF90_MAX_SLOTS=7
run LLMF90_PF_STAGES=0 LLMF90_SLOT_BYTES=16384 LLMF90_MAX_SLOTS=7
run LLMF90_PF_STAGES=16 LLMF90_SLOT_BYTES=8192 LLMF90_MAX_SLOTS=7
