#!/bin/bash
# refresh the single-GPU evidence: parity log, smoke, the four bench lines + reference arm, phase timings / traces,
# and the full ncu capture of the q4_0 kernel (gpurun_out/final_*)
mkdir -p gpurun_out
O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q --timeout=150 > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.log; tail -3 ${O}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > ${O}_bench_tinyllama_f32.json 2> ${O}_bench.err; echo "bench rc=$?"; cut -c1-200 ${O}_bench_tinyllama_f32.json
for cfg in "tinyllama f16" "llama2-7b q4_0" "llama2-7b f16"; do
  set -- $cfg
  timeout 500 python bench.py --steps 5 --warmup 3 --model $1 --wtype $2 > ${O}_bench_$1_$2.json 2>> ${O}_bench.err; cut -c1-200 ${O}_bench_$1_$2.json
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference_arm.json 2>> ${O}_bench.err; cut -c1-200 ${O}_bench_reference_arm.json
for cfg in "tinyllama f32" "tinyllama f16" "llama2-7b q4_0" "llama2-7b f16"; do
  set -- $cfg
  timeout 200 python tools/prof_phases.py $1 $2 > ${O}_phases_$1_$2.json 2>> ${O}_bench.err
  timeout 200 python tools/prof_trace.py $1 $2 10 64 > ${O}_trace_$1_$2.txt 2>> ${O}_bench.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stream_decode -s 60 -c 1 -f -o ${O}_prof_llama2-7b_q4_0 python tools/ncu_target.py llama2-7b q4_0 70 > ${O}_ncu_llama2-7b_q4_0.log 2>&1
tail -1 ${O}_ncu_llama2-7b_q4_0.log | cut -c1-200
