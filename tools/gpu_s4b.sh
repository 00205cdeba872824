#!/usr/bin/env bash
# A/B of the -D builds under tools/variants/ (two repetitions), then the parity tests of the step against the FASTEST
# build (its library copied over the in-tree one on the box's scratch copy of the repo) and against the default build
set -u
out=gpurun_out; mkdir -p $out; rm -f $out/variants.log
pkg=infantposeestimation_gaussianbias_b200
run() { LD_LIBRARY_PATH=$1 timeout 120 tools/bench_loss 1024 17 64 48 50 10 | sed "s/\"variant\": \"default\"/\"variant\": \"$2\"/" | tee -a $out/variants.log; }
for rep in 1 2; do
  run $pkg default
  for v in tools/variants/*/; do run $v $(basename $v); done
done
best=$(python - <<'P'
import json, collections
t = collections.defaultdict(list)
for l in open("gpurun_out/variants.log"):
    d = json.loads(l); t[d["variant"]].append(d["kernel_ms_mean"])
print(min(t, key=lambda k: sum(t[k]) / len(t[k])))
P
)
echo "fastest build: $best" | tee $out/best.txt
TESTS="tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_shapes.py tests/test_gpu_reference.py"
if [[ $best != default ]]; then
  cp $pkg/libgbcodec.so /tmp/libgbcodec.default.so
  cp tools/variants/$best/libgbcodec.so $pkg/libgbcodec.so
  timeout 1200 python -m pytest $TESTS -q -x > $out/pytest_best.log 2>&1; echo "pytest($best) rc=$?"; tail -3 $out/pytest_best.log
  # forward-only and no-variance calls of the same build
  for k in "5 3" ; do LD_LIBRARY_PATH=tools/variants/$best timeout 60 tools/bench_loss 256 17 64 48 $k; done
  cp /tmp/libgbcodec.default.so $pkg/libgbcodec.so
fi
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest(default) rc=$?"; tail -3 $out/pytest_gpu.log
du -sh $out
