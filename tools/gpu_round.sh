#!/bin/bash
# one GPU-box visit: parity tests, bench line, phase timings, per-CTA trace (outputs under gpurun_out/)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
for cfg in "tinyllama f32" "tinyllama f16" "tinyllama q4_0" "llama2-7b q4_0"; do
  set -- $cfg
  timeout 600 python tools/prof_phases.py $1 $2 > gpurun_out/phases_$1_$2.json 2> gpurun_out/phases_$1_$2.err
  cat gpurun_out/phases_$1_$2.json
done
timeout 300 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/trace_tinyllama_f32.txt 2>&1
cat gpurun_out/trace_tinyllama_f32.txt
