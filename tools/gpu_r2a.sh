#!/bin/bash
# round 2, call A: the new full-size parity tests on the round-1 kernel + a fresh per-CTA trace
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s > gpurun_out/r2a_pytest_fullsize.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_fullsize.log
tail -25 gpurun_out/r2a_pytest_fullsize.log
timeout 120 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2a_trace_tinyllama_f32.txt 2>&1
cat gpurun_out/r2a_trace_tinyllama_f32.txt
