#!/bin/bash
# Round-2f: ncu launch list of the default bench command (one whole 128-position generation after engine init).
set -u
mkdir -p gpurun_out
timeout -k 5 110 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 330 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_under_ncu.log 2>&1
echo "rc=$?" >> gpurun_out/r02f_bench_under_ncu.log
wc -l gpurun_out/r02f_launches_bench.csv
