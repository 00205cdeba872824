#!/usr/bin/env bash
# ncu --set full of one launch of every kernel family (workload: tools/bench_kernels.py --once --quick), summarised ON the
# box (the report is ~80 MB, more than gpurun brings back); plus the launch list of the default bench step.
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python tools/bench_kernels.py --once --quick > $out/once_plain.log 2>&1 || { tail -5 $out/once_plain.log; exit 1; }
K='regex:step_pipe_kernel|loss_tile_kernel|decode_tile_kernel|decode_warp_kernel|argmax_warp_kernel|encode_warp_kernel|encode_kernel|genb_tile_kernel|heatmap_step_kernel|postprocess_kernel|argmax_kernel|loss_tile_backward_kernel'
timeout 1500 ncu --set full --clock-control none -k "$K" -f -o /tmp/prof_all python tools/bench_kernels.py --once --quick > $out/ncu_all.log 2>&1
tail -2 $out/ncu_all.log
python tools/ncu_summary.py /tmp/prof_all.ncu-rep > $out/kernels_ncu_summary.txt 2>&1
wc -l $out/kernels_ncu_summary.txt
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras"
$CMD > $out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
tail -2 $out/ncu_launches.log
du -sh $out
