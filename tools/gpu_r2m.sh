#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2m_trace_tinyllama_f32.txt 2>&1; tail -9 gpurun_out/r2m_trace_tinyllama_f32.txt
timeout 150 python tools/prof_trace.py llama2-7b q4_0 10 64 > gpurun_out/r2m_trace_7b_q4.txt 2>&1; cat gpurun_out/r2m_trace_7b_q4.txt
