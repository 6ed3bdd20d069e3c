#!/bin/bash
mkdir -p gpurun_out
for m in "tinyllama f32 10 64" "tinyllama f32 10 128" "llama2-7b q4_0 10 64"; do
  set -- $m
  LLMF90_PF_LEAD=5 timeout 150 python tools/prof_trace.py $1 $2 $3 $4 > gpurun_out/r2f_trace_$1_$2_$4.txt 2>&1; tail -8 gpurun_out/r2f_trace_$1_$2_$4.txt
done
