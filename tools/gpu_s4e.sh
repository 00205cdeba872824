#!/usr/bin/env bash
# float16 instantiation of the persistent step kernel: its tests first (short time-outs: a pipeline that deadlocks must not
# hold the box), then timings (float16 step: default / two-deep ring / one-CTA-per-tile kernel; float32: default / V_EARLY=2),
# then the whole GPU suite
set -u
out=gpurun_out; mkdir -p $out; rm -f $out/f16_step.jsonl $out/variants.log
pkg=infantposeestimation_gaussianbias_b200
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "fp16" > $out/pytest_fp16.log 2>&1; rc=$?; echo "pytest fp16 rc=$rc"; tail -15 $out/pytest_fp16.log
if [[ $rc != 0 ]]; then echo "float16 step kernel failed its tests"; fi
timeout 200 python tools/bench_f16_step.py default | tee -a $out/f16_step.jsonl
cp $pkg/libgbcodec.so /tmp/libgbcodec.default.so
cp tools/variants/hring2/libgbcodec.so $pkg/libgbcodec.so
timeout 200 python tools/bench_f16_step.py "two-deep ring (-DPIPE_HALF_RING=2)" | tee -a $out/f16_step.jsonl
cp /tmp/libgbcodec.default.so $pkg/libgbcodec.so
run() { LD_LIBRARY_PATH=$1 timeout 120 tools/bench_loss 1024 17 64 48 50 10 | sed "s/\"variant\": \"default\"/\"variant\": \"$2\"/" | tee -a $out/variants.log; }
for rep in 1 2; do run $pkg default; run tools/variants/vearly2 vearly2; done
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest(all) rc=$?"; tail -4 $out/pytest_gpu.log
