#!/bin/bash
# ms/token of the device greedy loop for the bench configurations (no phase breakdown)
for cfg in "tinyllama f32" "tinyllama f16" "llama2-7b q4_0" "llama2-7b f16"; do
  set -- $cfg
  env "${@:3}" timeout 100 python tools/prof_phases.py $1 $2 --noprof 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['ms_per_token'],4))"
done
