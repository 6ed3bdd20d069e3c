#!/bin/bash
# round 2, call B: first run of the warp-owns-tile kernel.  Watchdog build first (a protocol bug ends in a
# printed diagnosis instead of a hang), then the production build: parity, bench, phases, trace.
mkdir -p gpurun_out
LLMF90_BUILD_WATCHDOG=1 python llm/f90_b200/build.py --force > gpurun_out/r2b_build_wd.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "transformer_matches_oracle or device_greedy" > gpurun_out/r2b_pytest_wd.log 2>&1; echo "wd pytest rc=$?"
grep -m5 WATCHDOG gpurun_out/r2b_pytest_wd.log; tail -5 gpurun_out/r2b_pytest_wd.log
if grep -q "passed" gpurun_out/r2b_pytest_wd.log && ! grep -q "failed\|WATCHDOG" gpurun_out/r2b_pytest_wd.log; then
  python llm/f90_b200/build.py --force > gpurun_out/r2b_build.log 2>&1
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
  tail -5 gpurun_out/r2b_pytest.log
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_tinyllama_f32.json 2> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench_tinyllama_f32.json | cut -c1-400
  for cfg in "tinyllama f32" "tinyllama f16" "llama2-7b q4_0" "llama2-7b f16"; do
    set -- $cfg
    timeout 200 python tools/prof_phases.py $1 $2 2> gpurun_out/r2b_phases_$1_$2.err | tee gpurun_out/r2b_phases_$1_$2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['ms_per_token'],4), {k: round(v,3) for k,v in d['phase_ms_per_token'].items()})" || tail -5 gpurun_out/r2b_phases_$1_$2.err
    timeout 200 python tools/prof_phases.py $1 $2 --noprof 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2 production ms/token', round(d['ms_per_token'],4))"
  done
  timeout 120 python tools/prof_trace.py tinyllama f32 10 64 > gpurun_out/r2b_trace_tinyllama_f32.txt 2>&1; cat gpurun_out/r2b_trace_tinyllama_f32.txt
fi
