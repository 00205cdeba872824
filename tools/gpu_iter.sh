#!/usr/bin/env bash
# one iteration on the GPU box: A/B builds of the step kernel, the GPU tests, the decode kernels warp-per-tile vs CTA-per-tile
set -u
out=gpurun_out; mkdir -p $out
bash tools/gpu_variants.sh
timeout 1500 python -m pytest tests -m gpu -q -s -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest_gpu.log
python tools/bench_decode.py --quick > $out/decode_warp.jsonl 2>$out/decode_warp.err; cat $out/decode_warp.jsonl
GBCODEC_DECODE_KERNEL=tile GBCODEC_ARGMAX_KERNEL=tile python tools/bench_decode.py --quick > $out/decode_tile.jsonl 2>$out/decode_tile.err; cat $out/decode_tile.jsonl
