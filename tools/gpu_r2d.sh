#!/bin/bash
mkdir -p gpurun_out
for m in "tinyllama f32" "tinyllama f16" "llama2-7b q4_0"; do
  set -- $m
  timeout 150 python tools/prof_trace.py $1 $2 10 64 > gpurun_out/r2d_trace_$1_$2.txt 2>&1; cat gpurun_out/r2d_trace_$1_$2.txt
done
