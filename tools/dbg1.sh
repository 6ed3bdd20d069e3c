timeout 100 python tools/prof_trace.py tinyllama f32 10 64 2>&1 | tail -6
