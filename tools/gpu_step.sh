#!/usr/bin/env bash
# One gpurun call for the persistent step kernel: small sanity runs (both kernels must print the same loss), timings of
# both kernels at the BASELINE size, then (optionally) part of the GPU test-suite.
#   gpurun --timeout 1500 -- 'bash tools/gpu_step.sh [pytest args...]'
set -u
out=gpurun_out; mkdir -p $out
for b in 2 8 64; do
  timeout 60 tools/bench_loss $b 17 64 48 2 1 || { echo "step kernel failed/hung at B=$b"; exit 1; }
  GBCODEC_STEP_KERNEL=tile timeout 60 tools/bench_loss $b 17 64 48 2 1
done
for i in 1 2 3; do
  timeout 120 tools/bench_loss 1024 17 64 48 50 10 | tee -a $out/step_time.log
done
GBCODEC_STEP_KERNEL=tile timeout 120 tools/bench_loss 1024 17 64 48 50 10 | tee -a $out/step_time.log
if [[ $# -gt 0 ]]; then
  timeout 1500 python -m pytest "$@" > $out/pytest_gpu.log 2>&1
  tail -15 $out/pytest_gpu.log
fi
