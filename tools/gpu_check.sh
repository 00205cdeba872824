#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, bench, launch list, one full ncu capture of the loss kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tests|bench|kernels|ncu|all]'
# Everything lands in gpurun_out/.
set -u
what="${1:-all}"
out=gpurun_out
mkdir -p $out
rc=0
if [[ $what == all || $what == tests ]]; then
  python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1 || rc=1
  tail -15 $out/pytest_gpu.log
  python __graft_entry__.py smoke > $out/smoke.log 2>&1 || rc=1
  tail -2 $out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  python bench.py > $out/bench.json 2> $out/bench.err || rc=1
  cat $out/bench.json
  python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2>> $out/bench.err || rc=1
  cat $out/bench_ref.json
fi
if [[ $what == all || $what == kernels ]]; then
  timeout 300 python tools/bench_kernels.py > $out/kernels.log 2>&1 || rc=1
  timeout 200 python tools/bench_decode.py > $out/decode.log 2>&1 || rc=1
  timeout 60 tools/bench_mix 17408 > $out/mix.log 2>&1 || rc=1
  tail -3 $out/kernels.log
fi
if [[ $what == all || $what == ncu ]]; then
  CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
  $CMD > $out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
  $CMD > $out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:loss_tile_kernel -s 3 -c 2 -f -o $out/prof_loss $CMD > $out/ncu_full.log 2>&1
  tail -3 $out/ncu_full.log
fi
exit $rc
