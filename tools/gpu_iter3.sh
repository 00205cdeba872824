#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
bash tools/gpu_variants.sh
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest_gpu.log
