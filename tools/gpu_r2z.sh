#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/prof_phases.py llama2-7b f16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('7b f16', round(d['ms_per_token'],4), {k: round(v,3) for k,v in d['phase_ms_per_token'].items()})"
timeout 200 python tools/prof_trace.py llama2-7b f16 10 64 > gpurun_out/r2z_trace_7b_f16.txt 2>&1; cat gpurun_out/r2z_trace_7b_f16.txt | cut -c1-250
