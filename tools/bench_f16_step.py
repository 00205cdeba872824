#!/usr/bin/env python
"""The fused step on float16 maps (autocast), B=1024 K=17 64x48: CUDA-event time per call of ops.fusion_loss_f16 for the
persistent step kernel (default) and the one-CTA-per-tile kernel (GBCODEC_STEP_F16=tile), the float32 step beside them.

    python tools/bench_f16_step.py [label]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import infantposeestimation_gaussianbias_b200 as pkg
pkg.load()
from infantposeestimation_gaussianbias_b200 import _native as N, ops

label = sys.argv[1] if len(sys.argv) > 1 else "default"
dev = torch.device("cuda", 0)
SK = ops.pairs_flat(((0, 1), (0, 2), (1, 3), (2, 4), (5, 6), (5, 7), (7, 9), (6, 8), (8, 10), (5, 11), (6, 12), (11, 12), (11, 13), (13, 15), (12, 14), (14, 16)))
LAM = [1.0, 1.0, 0.5, 0.1, 0.05, 0.05]
B, K, H, W = 1024, 17, 64, 48
g = torch.Generator(device=dev).manual_seed(0)
u = torch.rand(B, K, generator=g, device=dev)
vis = torch.where(u < 0.15, 0.0, torch.where(u < 0.40, 1.0, 2.0))
kps = torch.stack(((torch.rand(B, K, generator=g, device=dev) * 1.2 - 0.1) * 192, (torch.rand(B, K, generator=g, device=dev) * 1.2 - 0.1) * 256), -1).contiguous()
t, _ = ops.encode(kps + torch.randn(B, K, 2, generator=g, device=dev) * 6.0, torch.full_like(vis, 2.0), H, W, 192.0, 256.0, 2.0)
hm = t.mul_(torch.rand(B, K, 1, 1, generator=g, device=dev) * 0.9 + 0.3).add_(torch.randn(B, K, H, W, generator=g, device=dev), alpha=0.05)
off = torch.randn(B, K, 2, H, W, generator=g, device=dev).mul_(0.3)
var = torch.nn.functional.softplus(torch.randn(B, K, H, W, generator=g, device=dev))
h16 = dict(hm=hm.half(), off=off.half(), var=var.half())
alpha = torch.tensor([0.5], device=dev); fw = torch.tensor([0.6224593312018546], device=dev)
scale = torch.tensor([65536.0], device=dev)
DF = N.DECODE_REFINE | N.DECODE_APPLY_OFFSET


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


f16 = lambda grads=True: ops.fusion_loss_f16(h16["hm"], h16["off"], h16["var"], None, vis, kps, None, scale, 192.0, 256.0, LAM, 2.0, 2.0, True, SK,
                                             grads, True, alpha, fw, 2, DF)
f32 = lambda: ops.fusion_loss(hm, off, var, None, vis, kps, None, None, 192.0, 256.0, LAM, 2.0, 2.0, True, SK, True, True, alpha, fw, 2, DF)
n = H * W
rows = []
def report(what, ms, bytes_per_tile):
    r = dict(build=label, what=what, ms=round(ms, 4), heatmaps_per_s=round(B * K / (ms * 1e-3)), algorithmic_GBps=round(B * K * bytes_per_tile / (ms * 1e-3) / 1e9, 1))
    print(json.dumps(r), flush=True)
report("float16 step, persistent step kernel (12N bytes per tile)", timeit(f16), 12 * n)
report("float16 step, forward only", timeit(lambda: f16(False)), 4 * n)
os.environ["GBCODEC_STEP_F16"] = "tile"
report("float16 step, one-CTA-per-tile kernel (GBCODEC_STEP_F16=tile)", timeit(f16), 12 * n)
del os.environ["GBCODEC_STEP_F16"]
report("float32 step (24N bytes per tile)", timeit(f32), 24 * n)
