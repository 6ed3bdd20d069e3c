timeout 120 python tools/prof_trace.py tinyllama f16 10 64 2>&1 | tail -12
