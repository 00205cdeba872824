#!/usr/bin/env bash
# source-level ncu capture of loss_tile_kernel on the preemie shape (128x128, K=13) and on 96x72
set -u
out=gpurun_out; mkdir -p $out
tools/bench_loss 512 13 128 128 20 5 1.5 | tee $out/tile_128.log
tools/bench_loss 1024 17 96 72 20 5 | tee $out/tile_96.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:loss_tile_kernel -s 3 -c 1 -f -o $out/prof_tile128 tools/bench_loss 512 13 128 128 5 3 1.5 > $out/ncu_tile128.log 2>&1; tail -1 $out/ncu_tile128.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:loss_tile_kernel -s 3 -c 1 -f -o $out/prof_tile96 tools/bench_loss 1024 17 96 72 5 3 > $out/ncu_tile96.log 2>&1; tail -1 $out/ncu_tile96.log
