#!/bin/bash
# quick visit: the GPU parity suite (or a part of it: $1), then ms/token of the four bench configurations
mkdir -p gpurun_out
timeout 900 python -m pytest ${1:-tests} -m gpu -x -q --timeout=150 > gpurun_out/round_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/round_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
bash tools/ms_per_token.sh
