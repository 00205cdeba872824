#!/usr/bin/env bash
set -u
N=${1:-8}
out=gpurun_out; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 50 --warmup 5 ${N8_FLAGS:-} > $out/bench_n${N}.json 2> $out/bench_n${N}.err; echo "rc=$?"
python -c "
import json
d=json.loads(open('$out/bench_n${N}.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches_per_step'], d['config']['exchange'], d['total_loss'], (d['e2e'] or {}).get('value'))
print(d['other_workloads'])" || tail -20 $out/bench_n${N}.err
