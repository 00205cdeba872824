#!/usr/bin/env bash
# Decode-only workloads of BASELINE.json through bench.py (configs[2] and the configs[4] sweep).  GPU box only.
#   gpurun --timeout 900 -- 'bash tools/gpu_decode_bench.sh'
set -u
out=gpurun_out
mkdir -p $out
rc=0
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "host_buffer" > $out/pytest_host.log 2>&1 || rc=1
tail -3 $out/pytest_host.log
timeout 400 python bench.py --config decode_flip > $out/bench_decode_flip.json 2> $out/bench_decode.err || rc=1
timeout 300 python bench.py --config decode_flip --impl reference --steps 3 --warmup 1 > $out/bench_decode_flip_ref.json 2>> $out/bench_decode.err || rc=1
timeout 400 python bench.py --config decode > $out/bench_decode_16384.json 2>> $out/bench_decode.err || rc=1
for b in 256 1024 4096; do
  timeout 300 python bench.py --config decode --batch $b --no-cpu > $out/bench_decode_$b.json 2>> $out/bench_decode.err || rc=1
done
timeout 400 python bench.py --config decode --batch 65536 --no-cpu --no-e2e --steps 20 > $out/bench_decode_65536.json 2>> $out/bench_decode.err || rc=1
cat $out/bench_decode_flip.json $out/bench_decode_16384.json
tail -5 $out/bench_decode.err
exit $rc
