#!/usr/bin/env bash
# round-2 first call: sanity of both kernels, timing, ncu full capture of the persistent step kernel, GPU test-suite
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt
for b in 2 64; do
  timeout 60 tools/bench_loss $b 17 64 48 2 1 || { echo "step kernel failed/hung at B=$b"; exit 1; }
  GBCODEC_STEP_KERNEL=tile timeout 60 tools/bench_loss $b 17 64 48 2 1
done
for i in 1 2; do
  timeout 120 tools/bench_loss 1024 17 64 48 50 10 | tee -a $out/step_time.log
  GBCODEC_STEP_KERNEL=tile timeout 120 tools/bench_loss 1024 17 64 48 50 10 | tee -a $out/step_time.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_tile_kernel -s 3 -c 1 -f -o $out/prof_step tools/bench_loss 1024 17 64 48 5 3 > $out/ncu_step.log 2>&1
tail -2 $out/ncu_step.log
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1
tail -8 $out/pytest_gpu.log
