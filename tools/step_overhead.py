#!/usr/bin/env python
"""Where does the host time of one resident fused step go?  (measurement aid, GPU box only)"""
import cProfile, pstats, io, os, sys, time
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

def timed(n, label, sampler=False):
    s = None
    if sampler:
        s = bench.ClockSampler(0); s.start(); time.sleep(0.3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n):
        res = step()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    if s: s.stop()
    print(f"{label}: device {a.elapsed_time(b)/n:.3f} ms/step, host issue {(t1-t0)/n*1e3:.3f} ms/step, wall {(t2-t0)/n*1e3:.3f}")

for _ in range(5): step()
timed(50, "plain")
timed(50, "plain again")
timed(50, "with nvidia-smi sampler", sampler=True)
timed(50, "after sampler")
print(torch.cuda.memory_stats()["num_device_alloc"], "device allocs so far")
pr = cProfile.Profile(); pr.enable()
for _ in range(50): res = step()
torch.cuda.synchronize(); pr.disable()
print(torch.cuda.memory_stats()["num_device_alloc"], "device allocs after")
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(25); print(st.getvalue()[:6000])
