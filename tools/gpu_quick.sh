#!/bin/bash
# quick GPU visit: parity tests + phase timings of the four bench configs + per-CTA trace
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for cfg in "tinyllama f32" "tinyllama f16" "tinyllama q4_0" "llama2-7b q4_0" "llama2-7b f16"; do
  set -- $cfg
  timeout 100 python tools/prof_phases.py $1 $2 2> gpurun_out/phases_$1_$2.err | tee gpurun_out/phases_$1_$2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['ms_per_token'],4), {k: round(v,3) for k,v in d['phase_ms_per_token'].items()})" || tail -5 gpurun_out/phases_$1_$2.err
done
timeout 100 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/trace_tinyllama_f32.txt 2>&1
cat gpurun_out/trace_tinyllama_f32.txt
