#!/bin/bash
# Round-2d: what the driver runs at round end, on the final tree: the whole GPU suite, smoke(), the default bench line.
set -u
mkdir -p gpurun_out
timeout -k 5 260 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02d_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest_gpu.log
timeout -k 5 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02d_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02d_smoke.log
timeout -k 5 150 python bench.py > gpurun_out/r02d_bench_default.json 2> gpurun_out/r02d_bench_err.txt
echo "bench rc=$?" >> gpurun_out/r02d_bench_err.txt
tail -4 gpurun_out/r02d_pytest_gpu.log; tail -3 gpurun_out/r02d_smoke.log; cut -c1-300 gpurun_out/r02d_bench_default.json
