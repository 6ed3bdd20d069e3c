#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2n_trace_probe.txt 2>&1; tail -5 gpurun_out/r2n_trace_probe.txt
timeout 150 python tools/prof_trace.py llama2-7b q4_0 10 64 > gpurun_out/r2n_trace_probe_q4.txt 2>&1; tail -5 gpurun_out/r2n_trace_probe_q4.txt
