#!/usr/bin/env bash
# final evidence bundle at N=1: smoke, default bench line (with extras), reference arm, whole GPU test-suite with its
# reports, ncu --set full of the step kernel (source-correlated) and of the warp-per-tile decode / arg-max kernels,
# the three step-kernel designs on one box
set -u
out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
tools/bench_loss 1024 17 64 48 5 3 > $out/plain_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_pipe_kernel -s 3 -c 1 -f -o $out/prof_step tools/bench_loss 1024 17 64 48 5 3 > $out/ncu_step.log 2>&1; tail -1 $out/ncu_step.log
if [[ "${FINAL_DECODE_NCU:-0}" == 1 ]]; then
python tools/decode_once.py > $out/decode_once.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -k 'regex:decode_warp_kernel|argmax_warp_kernel' -s 4 -c 2 -f -o /tmp/prof_decode python tools/decode_once.py > $out/ncu_decode.log 2>&1
python tools/ncu_summary.py /tmp/prof_decode.ncu-rep > $out/decode_warp_ncu_summary.txt 2>&1; wc -l $out/decode_warp_ncu_summary.txt
fi
rm -f $out/designs.log
for k in pipe persist tile; do GBCODEC_STEP_KERNEL=$k timeout 120 tools/bench_loss 1024 17 64 48 50 10 | sed "s/\"variant\": \"default\"/\"variant\": \"$k\"/" | tee -a $out/designs.log; done
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras"
$CMD > $out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
tail -2 $out/ncu_launches.log
du -sh $out
