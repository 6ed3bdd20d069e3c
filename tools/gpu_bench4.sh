#!/bin/bash
# refresh the four single-GPU bench lines (profiles/r02_bench_*.json)
mkdir -p gpurun_out
O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q --timeout=150 > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.log; tail -3 ${O}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > ${O}_bench_tinyllama_f32.json 2> ${O}_bench.err; echo "bench rc=$?"; cut -c1-200 ${O}_bench_tinyllama_f32.json
for cfg in "tinyllama f16" "llama2-7b q4_0" "llama2-7b f16"; do
  set -- $cfg
  timeout 500 python bench.py --steps 5 --warmup 3 --model $1 --wtype $2 > ${O}_bench_$1_$2.json 2>> ${O}_bench.err; cut -c1-200 ${O}_bench_$1_$2.json
done
for cfg in "tinyllama f32" "tinyllama f16" "llama2-7b q4_0" "llama2-7b f16"; do
  set -- $cfg
  timeout 200 python tools/prof_phases.py $1 $2 > ${O}_phases_$1_$2.json 2>> ${O}_bench.err
  timeout 200 python tools/prof_trace.py $1 $2 10 64 > ${O}_trace_$1_$2.txt 2>> ${O}_bench.err
done
