#!/bin/bash
# 8-GPU visit: memory check first (every rank builds the full synthetic model on the host), then the TP bench line
mkdir -p gpurun_out
free -g | head -2; nproc
MEM=$(free -g | awk '/Mem:/{print $7}')
if [ "$MEM" -lt 150 ]; then echo "LOWMEM: $MEM GB available, skipping the 8-rank bench"; exit 0; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_tp8.json 2> gpurun_out/bench_tp8.err
echo "rc=$?"; tail -1 gpurun_out/bench_tp8.json | cut -c1-1800; tail -3 gpurun_out/bench_tp8.err
