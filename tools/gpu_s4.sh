#!/usr/bin/env bash
# session 4 of round 2, one gpurun call: A/B of the step kernel's -D builds (tools/variants/), the whole GPU suite,
# the per-kernel bench (float16 Gen-B line), ncu --set full of the step kernel as it is now
set -u
out=gpurun_out; mkdir -p $out; rm -f $out/variants.log
run() { LD_LIBRARY_PATH=$1 timeout 120 tools/bench_loss 1024 17 64 48 50 10 | sed "s/\"variant\": \"default\"/\"variant\": \"$2\"/" | tee -a $out/variants.log; }
for rep in 1 2; do
  run infantposeestimation_gaussianbias_b200 default
  for v in tools/variants/*/; do run $v $(basename $v); done
done
timeout 1500 python -m pytest tests -m gpu -q -s -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.log
timeout 600 python tools/bench_kernels.py --quick > $out/kernels_quick.jsonl 2> $out/kernels_quick.err; echo "bench_kernels rc=$?"; grep -i "genb\|CombinedLoss" $out/kernels_quick.jsonl | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_pipe_kernel -s 3 -c 1 -f -o $out/prof_step tools/bench_loss 1024 17 64 48 5 3 > $out/ncu_step.log 2>&1; tail -1 $out/ncu_step.log
timeout 300 python bench.py --no-e2e --no-cpu --no-extras > $out/bench_quick.json 2> $out/bench_quick.err; echo "bench rc=$?"; cut -c1-400 $out/bench_quick.json
du -sh $out
