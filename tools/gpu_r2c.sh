#!/bin/bash
# round 2, call C: warp-owns-tile kernel with the pipelined tile_dot: parity, phases, trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c_pytest.log
for cfg in "tinyllama f32" "tinyllama f16" "tinyllama q4_0" "llama2-7b q4_0" "llama2-7b f16"; do
  set -- $cfg
  timeout 200 python tools/prof_phases.py $1 $2 2> gpurun_out/r2c_phases_$1_$2.err | tee gpurun_out/r2c_phases_$1_$2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['ms_per_token'],4), {k: round(v,3) for k,v in d['phase_ms_per_token'].items()})" || tail -5 gpurun_out/r2c_phases_$1_$2.err
  timeout 200 python tools/prof_phases.py $1 $2 --noprof 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 production ms/token', round(d['ms_per_token'],4))"
done
timeout 120 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2c_trace_tinyllama_f32.txt 2>&1; cat gpurun_out/r2c_trace_tinyllama_f32.txt
