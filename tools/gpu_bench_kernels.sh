#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
for k in pipe tile persist pipe; do
  GBCODEC_STEP_KERNEL=$k timeout 300 python bench.py --no-e2e --no-cpu --no-extras --steps 50 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$k', d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches_per_step'], d['clocks'])"
done
timeout 120 tools/bench_loss 1024 17 64 48 50 10
