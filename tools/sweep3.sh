#!/bin/bash
run() { M=$1; W=$2; shift 2; echo "== $M $W $*"; env "$@" timeout 100 python tools/prof_phases.py $M $W 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_token'],4), {k: round(v,3) for k,v in d['phase_ms_per_token'].items()})"; }
run llama2-7b q4_0 A=0
run tinyllama q4_0 A=0
