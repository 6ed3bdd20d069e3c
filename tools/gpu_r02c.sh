#!/bin/bash
# Round-2c: the batched prompt pass at full size, the benchmark line with and without it, the new GPU test files,
# one ncu --set full capture of the tcgen05 GEMM and a launch list of a generation with the prompt pass.
set -u
mkdir -p gpurun_out
export LLMF90_WORKER_NO_BUILD=1
W="python tests/prefill_worker.py"
timeout -k 5 120 $W multi "prefill mid 0 130" "prefill tinyllama 0 16" "prefill tinyllama 2 16" > gpurun_out/r02c_cases.jsonl 2> gpurun_out/r02c_err_cases.txt
echo "{\"group\": \"cases\", \"rc\": $?}" >> gpurun_out/r02c_cases.jsonl
timeout -k 5 150 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench_tinyllama_f32.json 2> gpurun_out/r02c_bench_err.txt
echo "bench rc $?" >> gpurun_out/r02c_bench_err.txt
timeout -k 5 240 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_q6k.py tests/test_gpu_sampler.py -q > gpurun_out/r02c_pytest_new.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02c_pytest_new.log
timeout -k 5 100 python bench.py --steps 5 --warmup 3 --no-prefill --no-cpu-baseline > gpurun_out/r02c_bench_tinyllama_f32_no_prefill.json 2>> gpurun_out/r02c_bench_err.txt
timeout -k 5 90 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -c 1 -f -o gpurun_out/r02c_umma_w13 $W matmul 0 11264 2048 16 > gpurun_out/r02c_ncu_full.log 2>&1
timeout -k 5 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches.csv $W greedy small 0 9 40 > gpurun_out/r02c_ncu_list.log 2>&1
timeout -k 5 120 python bench.py --steps 5 --warmup 3 --model llama2-7b --wtype q4_0 --no-cpu-baseline > gpurun_out/r02c_bench_llama2_7b_q4_0.json 2>> gpurun_out/r02c_bench_err.txt
cat gpurun_out/r02c_cases.jsonl; tail -3 gpurun_out/r02c_pytest_new.log; cut -c1-600 gpurun_out/r02c_bench_tinyllama_f32.json
