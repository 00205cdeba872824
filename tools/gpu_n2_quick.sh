#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p $out
for ex in peer peer-sync; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 5 --exchange $ex --no-e2e --no-extras > $out/q.json 2> $out/q.err
python -c "
import json
d=json.loads(open('$out/q.json').read().strip().splitlines()[-1])
print('$ex', d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches_per_step'], 'host', d['host_enqueue_ms_per_step'])" || tail -5 $out/q.err
done
timeout 300 python bench.py --no-e2e --no-cpu --no-extras | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('n1', d['ms_per_step'], d['roofline']['kernel_ms'], 'host', d['host_enqueue_ms_per_step'])"
