#!/usr/bin/env bash
# final evidence bundle at N=1: smoke, default bench line (with extras), reference arm, whole GPU test-suite with its
# reports, ncu --set full of the step kernel (source-correlated)
set -u
out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
tools/bench_loss 1024 17 64 48 5 3 > $out/plain_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_pipe_kernel -s 3 -c 1 -f -o $out/prof_step tools/bench_loss 1024 17 64 48 5 3 > $out/ncu_step.log 2>&1; tail -1 $out/ncu_step.log
for k in pipe persist tile; do GBCODEC_STEP_KERNEL=$k timeout 120 tools/bench_loss 1024 17 64 48 50 10 | sed "s/\"variant\": \"default\"/\"variant\": \"$k\"/" | tee -a $out/designs.log; done
du -sh $out
