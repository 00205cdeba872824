#!/usr/bin/env bash
# ncu captures of the persistent step kernel (one gpurun call): a quick section list first, then --set full
set -u
out=gpurun_out; mkdir -p $out
tools/bench_loss 1024 17 64 48 5 3 > $out/plain_step.log 2>&1 || { cat $out/plain_step.log; exit 1; }
cat $out/plain_step.log
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section InstructionStats --section ComputeWorkloadAnalysis"
timeout 600 ncu $SECS --clock-control none -k regex:step_pipe_kernel -s 3 -c 1 -f -o $out/prof_step_quick tools/bench_loss 1024 17 64 48 5 3 > $out/ncu_step_quick.log 2>&1
tail -3 $out/ncu_step_quick.log
timeout ${1:-1500} ncu --set full --clock-control none --import-source on -k regex:step_pipe_kernel -s 3 -c 1 -f -o $out/prof_step tools/bench_loss 1024 17 64 48 5 3 > $out/ncu_step.log 2>&1
tail -3 $out/ncu_step.log
ls -la $out
