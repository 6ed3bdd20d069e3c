#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2v_trace_tinyllama_f32.txt 2>&1; cat gpurun_out/r2v_trace_tinyllama_f32.txt | cut -c1-260
bash tools/ms_per_token.sh
