#!/usr/bin/env python
"""Host time of each of the first resident steps, queued without synchronising (measurement aid, GPU box only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import infantposeestimation_gaussianbias_b200 as pkg
pkg.load()
from infantposeestimation_gaussianbias_b200 import _native as N, ops
dev = torch.device("cuda", 0)
B = 1024
pairs = ops.pairs_flat(bench.SKELETON)
data = bench.synth_device_batch(B, dev, 1234, ops)
alpha = torch.tensor([0.5], device=dev); fw = torch.tensor([0.62], device=dev)
dflags = N.DECODE_REFINE | N.DECODE_APPLY_OFFSET
def step():
    return ops.fusion_loss(data["hm"], data["off"], data["var"], None, data["vis"], data["kps"], None, None,
                           float(bench.IN_W), float(bench.IN_H), bench.LAMBDAS, bench.SIGMA, bench.SIGMA, True, pairs, True, True, alpha, fw, 2, dflags)
for _ in range(5): step()
torch.cuda.synchronize()
ts = []
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t_prev = time.perf_counter()
if len(sys.argv) > 1: a.record()
print("record", (time.perf_counter() - t_prev) * 1e3)
for i in range(60):
    res = step()
    t = time.perf_counter()
    ts.append(((t - t_prev) * 1e3, 0, 0))
    t_prev = t
torch.cuda.synchronize()
print("sync", (time.perf_counter() - t_prev) * 1e3)
for i, t in enumerate(ts):
    if t[0] > 0.3 or i < 8: print(i, f"issue {t[0]:.3f} ms allocs {t[1]} reserved {t[2]} MiB")
