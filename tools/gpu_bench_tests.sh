#!/usr/bin/env bash
# default bench line + reference arm + GPU tests (one gpurun call)
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"; tail -3 $out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep "real-reference" $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
