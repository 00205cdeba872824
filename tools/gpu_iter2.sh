#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
bash tools/gpu_variants.sh
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
