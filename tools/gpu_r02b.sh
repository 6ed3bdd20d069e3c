#!/bin/bash
# Round-2b hardware check of the code written without a GPU (one short call: the round's GPU budget was nearly spent):
# the tcgen05 GEMM (with the descriptor variants whose reading of the ISA could not be tested here), the batched prompt
# pass, the Q6_K classifier and the device sampler.  Every group runs in its own process under a timeout and appends
# one JSON line per case to gpurun_out/r02b_cases.jsonl.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r02b_cases.jsonl
: > $OUT
export LLMF90_WORKER_NO_BUILD=1
W="python tests/prefill_worker.py"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r02b_gpu.txt 2>&1
# 1. proven-kernel neighbours first: Q6_K and the sampler (plain CUDA)
timeout -k 5 150 $W multi "q6k_matvec 37 512" "q6k_matvec 1000 2048" "q6k_model 2" "q6k_model 0" "sample small 0" >> $OUT 2> gpurun_out/r02b_err_1.txt
echo "{\"group\": 1, \"rc\": $?}" >> $OUT
# 2. the tcgen05 GEMM, smallest case, descriptor variants 0..3
GOOD=""
for v in 0 1 2 3; do
  LLMF90_UMMA_SWAP_LBO=$v timeout -k 5 60 $W multi "matmul 1 128 64 16" "matmul 0 37 96 5" > gpurun_out/r02b_v$v.txt 2> gpurun_out/r02b_err_v$v.txt
  rc=$?
  sed "s/^{/{\"variant\": $v, /" gpurun_out/r02b_v$v.txt >> $OUT
  echo "{\"group\": \"variant $v\", \"rc\": $rc}" >> $OUT
  if [ -z "$GOOD" ] && python - gpurun_out/r02b_v$v.txt <<'PY'
import json, sys
ok = 0
for ln in open(sys.argv[1]):
    try:
        d = json.loads(ln)
    except Exception:
        continue
    if d.get("rel_err", 1) < 1e-3:
        ok += 1
sys.exit(0 if ok == 2 else 1)
PY
  then GOOD=$v; fi
done
echo "{\"good_variant\": \"$GOOD\"}" >> $OUT
export LLMF90_UMMA_SWAP_LBO=${GOOD:-0}
# 3. larger GEMMs and the batched prompt pass with the variant that worked
timeout -k 5 120 $W multi "matmul 0 2560 2048 16" "matmul 2 1000 4096 128" "matmul 1 2048 5632 33" >> $OUT 2> gpurun_out/r02b_err_3.txt
echo "{\"group\": 3, \"rc\": $?}" >> $OUT
timeout -k 5 150 $W multi "prefill tiny 0 5" "prefill small 1 16" "prefill mha 2 37" "prefill mid 0 130" "greedy small 0 9 40" "greedy small 2 9 40" >> $OUT 2> gpurun_out/r02b_err_4.txt
echo "{\"group\": 4, \"rc\": $?}" >> $OUT
# 4. the CLI with the device pick (argmax path of the verified tests) and a Q6_K file, through pytest if time is left
timeout -k 5 200 python -m pytest tests/test_gpu_q6k.py tests/test_gpu_parity.py -x -q -k "cli or q6" > gpurun_out/r02b_pytest_cli.txt 2>&1
echo "{\"group\": 5, \"rc\": $?}" >> $OUT
cat $OUT
