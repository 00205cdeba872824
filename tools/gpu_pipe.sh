#!/usr/bin/env bash
# sanity of the pipelined step kernel against the other two (same loss), timings, then GPU tests
set -u
out=gpurun_out; mkdir -p $out
for b in 1 2 8 64 300; do
  for k in pipe persist tile; do
    GBCODEC_STEP_KERNEL=$k timeout 60 tools/bench_loss $b 17 64 48 3 1 | sed "s/\"variant\": \"default\"/\"variant\": \"$k\"/" || { echo "$k failed/hung at B=$b"; [[ $k == pipe ]] && exit 1; }
  done
done
for rep in 1 2; do
  for k in pipe persist tile; do
    GBCODEC_STEP_KERNEL=$k timeout 120 tools/bench_loss 1024 17 64 48 50 10 | sed "s/\"variant\": \"default\"/\"variant\": \"$k\"/" | tee -a $out/pipe_time.log
  done
done
if [[ $# -gt 0 ]]; then
  timeout 1500 python -m pytest "$@" > $out/pytest_gpu.log 2>&1
  tail -40 $out/pytest_gpu.log
fi
