#!/usr/bin/env bash
# ncu --set full of one launch of every kernel family (workload: tools/bench_kernels.py --once --quick), plus the launch
# list of the default bench step.  Summaries: python tools/ncu_summary.py gpurun_out/prof_all.ncu-rep
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python tools/bench_kernels.py --once --quick > $out/once_plain.log 2>&1 || { tail -5 $out/once_plain.log; exit 1; }
K='regex:step_pipe_kernel|loss_tile_kernel|decode_tile_kernel|encode_warp_kernel|encode_kernel|genb_tile_kernel|heatmap_step_kernel|postprocess_kernel|argmax_kernel|loss_tile_backward_kernel'
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" -f -o $out/prof_all python tools/bench_kernels.py --once --quick > $out/ncu_all.log 2>&1
tail -3 $out/ncu_all.log
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras"
$CMD > $out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
tail -2 $out/ncu_launches.log
ls -la $out | tail -8
