#!/usr/bin/env bash
# One gpurun call for the persistent step kernel: small sanity run (both kernels must print the same loss), the GPU
# test-suite, then timings of both kernels at the BASELINE sizes.
#   gpurun --timeout 1500 -- 'bash tools/gpu_step.sh [quick|tests|time|all]'
set -u
what="${1:-all}"
out=gpurun_out
mkdir -p $out
rc=0
if [[ $what == all || $what == quick ]]; then
  for b in 2 8 64; do
    timeout 60 tools/bench_loss $b 17 64 48 2 1 || { echo "step kernel failed/hung at B=$b"; exit 1; }
    GBCODEC_STEP_KERNEL=tile timeout 60 tools/bench_loss $b 17 64 48 2 1 || rc=1
  done
fi
if [[ $what == all || $what == tests ]]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1 || rc=1
  tail -15 $out/pytest_gpu.log
fi
if [[ $what == all || $what == time ]]; then
  for i in 1 2; do
    timeout 120 tools/bench_loss 1024 17 64 48 50 10 | tee -a $out/step_time.log
    GBCODEC_STEP_KERNEL=tile timeout 120 tools/bench_loss 1024 17 64 48 50 10 | tee -a $out/step_time.log
  done
fi
exit $rc
