#!/bin/bash
# Round-2e: last check of the final tree: the prompt-pass tests (incl. `llm --prefill`), the sampler / Q6_K files, a short default bench.
set -u
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_q6k.py tests/test_gpu_sampler.py -x -q > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
timeout -k 5 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench_err.txt
echo "bench rc=$?" >> gpurun_out/r02e_bench_err.txt
tail -3 gpurun_out/r02e_pytest.log; cut -c1-200 gpurun_out/r02e_bench.json
