#!/usr/bin/env bash
# what the decode costs inside the step kernel: the default library with and without the decode outputs
set -u
out=gpurun_out; mkdir -p $out
for rep in 1 2; do
  timeout 120 tools/bench_loss 1024 17 64 48 50 10 | tee -a $out/nodecode.log
  BENCH_NODECODE=1 timeout 120 tools/bench_loss 1024 17 64 48 50 10 | sed 's/"variant": "default"/"variant": "no decode"/' | tee -a $out/nodecode.log
done
