#!/bin/bash
# one ncu --set full capture of the fused decode kernel (position ~61 of a TinyLlama f32 run)
mkdir -p gpurun_out
M=${1:-tinyllama}; W=${2:-f32}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stream_decode -s 60 -c 1 -f -o gpurun_out/prof_${M}_${W} python tools/ncu_target.py $M $W 70 > gpurun_out/ncu_${M}_${W}.log 2>&1
tail -3 gpurun_out/ncu_${M}_${W}.log
