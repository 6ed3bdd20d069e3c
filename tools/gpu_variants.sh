#!/bin/bash
# A/B test of prebuilt library variants (variants/lib*.so)
mkdir -p gpurun_out
cp llm/f90_b200/libllmf90_b200.so /tmp/lib_orig.so
for v in variants/lib*.so; do
  cp $v llm/f90_b200/libllmf90_b200.so
  echo "== $v"
  bash tools/ms_per_token.sh
done
cp /tmp/lib_orig.so llm/f90_b200/libllmf90_b200.so
