#!/bin/bash
# round-end evidence run on one B200: parity tests, smoke, bench lines (ours + reference arm), phase timings and
# a layer trace, the ncu launch list and full captures of the three kernels; everything lands in gpurun_out/final_*
mkdir -p gpurun_out
O=gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q --timeout=150 > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.log; tail -3 ${O}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > ${O}_bench_tinyllama_f32.json 2> ${O}_bench.err; echo "bench rc=$?"; cat ${O}_bench_tinyllama_f32.json | cut -c1-2500
for cfg in "tinyllama f16" "llama2-7b q4_0" "llama2-7b f16"; do
  set -- $cfg
  timeout 500 python bench.py --steps 5 --warmup 3 --model $1 --wtype $2 > ${O}_bench_$1_$2.json 2>> ${O}_bench.err; cat ${O}_bench_$1_$2.json | cut -c1-300
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference_arm.json 2>> ${O}_bench.err; cat ${O}_bench_reference_arm.json | cut -c1-600
for cfg in "tinyllama f32" "tinyllama f16" "llama2-7b q4_0" "llama2-7b f16"; do
  set -- $cfg
  timeout 200 python tools/prof_phases.py $1 $2 > ${O}_phases_$1_$2.json 2>> ${O}_bench.err
  timeout 200 python tools/prof_trace.py $1 $2 10 64 > ${O}_trace_$1_$2.txt 2>> ${O}_bench.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${O}_launches_bench.log 2>&1
tail -2 ${O}_launches.csv
for cfg in "tinyllama f32" "tinyllama f16" "llama2-7b q4_0"; do
  set -- $cfg
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:stream_decode -s 60 -c 1 -f -o ${O}_prof_$1_$2 python tools/ncu_target.py $1 $2 70 > ${O}_ncu_$1_$2.log 2>&1
  tail -1 ${O}_ncu_$1_$2.log | cut -c1-200
done
ls -la gpurun_out/final_* | awk '{print $5, $9}'
