#!/usr/bin/env bash
# time tools/bench_loss against every library under tools/variants/ and the in-tree one; then the GPU tests
set -u
out=gpurun_out; mkdir -p $out
run() { LD_LIBRARY_PATH=$1 timeout 120 tools/bench_loss 1024 17 64 48 50 10 | sed "s/\"variant\": \"default\"/\"variant\": \"$2\"/" | tee -a $out/variants.log; }
for rep in 1 2; do
  run infantposeestimation_gaussianbias_b200 default
  for v in tools/variants/*/; do run $v $(basename $v); done
done
if [[ $# -gt 0 ]]; then
  timeout 1500 python -m pytest "$@" > $out/pytest_gpu.log 2>&1
  tail -40 $out/pytest_gpu.log
fi
