#!/usr/bin/env bash
# N-rank runs (gpurun --gpus N): sharded tests, then bench.py under torchrun with the three exchange modes
set -u
N=${1:-2}
out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -8
if [[ "${2:-}" == tests ]]; then
  timeout 900 python -m pytest tests/test_gpu_sharded_peer.py -m gpu -q -x > $out/pytest_sharded_n$N.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest_sharded_n$N.log
fi
port=29500
for ex in peer peer-sync nccl; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 50 --warmup 5 --exchange $ex --no-e2e --no-extras > $out/bench_n${N}_$ex.json 2> $out/bench_n${N}_$ex.err
  echo "$ex rc=$?"
  python -c "
import json,sys
d=json.loads(open('$out/bench_n${N}_$ex.json').read().strip().splitlines()[-1])
print('$ex', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches_per_step'], d['config']['exchange'], d['total_loss'])" || tail -5 $out/bench_n${N}_$ex.err
done
