#!/bin/bash
# 8-GPU visit: TP parity tests at world 2 / 4 / 8 (incl. replicated KV heads), then the TP bench lines at N = 4 and 8
mkdir -p gpurun_out
free -g | head -2; nproc
MEM=$(free -g | awk '/Mem:/{print $7}')
timeout 500 python -m pytest tests/test_gpu_tp.py -m gpu -x -v --timeout=150 > gpurun_out/tp8_pytest.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/tp8_pytest.log
if [ "$MEM" -lt 150 ]; then echo "LOWMEM: $MEM GB available, skipping the 8-rank bench"; exit 0; fi
for N in 4 8; do
  EXTRA="--no-tp1"; [ $N = 8 ] && EXTRA=""
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 3 --warmup 3 $EXTRA > gpurun_out/bench_tp$N.json 2> gpurun_out/bench_tp$N.err
  echo "N=$N rc=$?"; tail -1 gpurun_out/bench_tp$N.json | cut -c1-400; tail -2 gpurun_out/bench_tp$N.err | cut -c1-300
done
