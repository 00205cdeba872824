#!/usr/bin/env bash
# validation of the last changes: whole GPU suite, the float32 step's time (must not move), per-kernel bench (per-tile-mean
# step through the persistent kernel), default bench line (api_step after the host-side fast paths)
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.log
for rep in 1 2; do timeout 120 tools/bench_loss 1024 17 64 48 50 10; done | tee $out/step_after_vmean.log
timeout 600 python tools/bench_kernels.py --quick > $out/kernels_quick.jsonl 2> $out/kernels_quick.err; echo "bench_kernels rc=$?"
grep -i "variance means\|loss step\|f16\|float16" $out/kernels_quick.jsonl | cut -c1-260
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"])
print({k: (v["ms_per_step"], v["host_enqueue_ms_per_step"]) for k, v in d["api_step"].items() if isinstance(v, dict)})
P
