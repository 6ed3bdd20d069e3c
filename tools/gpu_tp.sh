#!/bin/bash
# multi-GPU visit: TP parity tests + the TP bench line (N = number of GPUs of this box)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_raw.txt 2>&1; head -8 gpurun_out/nvlink_raw.txt
if [ -z "$SKIP_TESTS" ]; then
timeout 400 python -m pytest tests/test_gpu_tp.py -m gpu -x -q --timeout=150 > gpurun_out/tp${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tp${N}_pytest.log
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 $BENCH_EXTRA > gpurun_out/bench_tp$N.json 2> gpurun_out/bench_tp$N.err
tail -2 gpurun_out/bench_tp$N.json | cut -c1-300
tail -3 gpurun_out/bench_tp$N.err | cut -c1-300
