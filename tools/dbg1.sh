timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for cfg in "tinyllama q4_0" "llama2-7b q4_0"; do
  set -- $cfg
  timeout 100 python tools/prof_phases.py $1 $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['ms_per_token'],4), {k: round(v,3) for k,v in d['phase_ms_per_token'].items()})"
done
