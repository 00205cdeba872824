#!/usr/bin/env python
"""Roofline of every codec kernel on the BASELINE.json configurations (GPU box only).

    python tools/bench_kernels.py [--quick] > gpurun_out/kernels.json

Each case: inputs resident in HBM and larger than L2, 3 warm-up + N timed launches, CUDA events on the
launching stream, algorithmic bytes per tile from DESIGN.md §4, peak = MEASURED_PEAKS.json hbm_gbs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import infantposeestimation_gaussianbias_b200 as pkg
pkg.load()
from infantposeestimation_gaussianbias_b200 import _native as N, ops

dev = torch.device("cuda", 0)
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
SK = ops.pairs_flat(((0, 1), (0, 2), (1, 3), (2, 4), (5, 6), (5, 7), (7, 9), (6, 8), (8, 10), (5, 11), (6, 12), (11, 12), (11, 13), (13, 15), (12, 14), (14, 16)))
LAM = [1.0, 1.0, 0.5, 0.1, 0.05, 0.05]
quick = "--quick" in sys.argv
once = "--once" in sys.argv          # one launch per case, untimed: the workload of the ncu captures (tools/gpu_ncu_all.sh)


def timeit(fn, n=20):
    if once:
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(3):
        r = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def inputs(B, K, H, W, in_w, in_h, sigma, seed=0, flip=False, offsets=True, var=False):
    g = torch.Generator(device=dev).manual_seed(seed)
    u = torch.rand(B, K, generator=g, device=dev)
    vis = torch.where(u < 0.15, 0.0, torch.where(u < 0.40, 1.0, 2.0))
    kps = torch.stack(((torch.rand(B, K, generator=g, device=dev) * 1.2 - 0.1) * in_w,
                       (torch.rand(B, K, generator=g, device=dev) * 1.2 - 0.1) * in_h), -1).contiguous()
    t, _ = ops.encode(kps + torch.randn(B, K, 2, generator=g, device=dev) * 6.0, torch.full_like(vis, 2.0), H, W, float(in_w), float(in_h), sigma)
    hm = t.mul_(torch.rand(B, K, 1, 1, generator=g, device=dev) * 0.9 + 0.3)
    hm.add_(torch.randn(B, K, H, W, generator=g, device=dev), alpha=0.05)
    d = dict(kps=kps, vis=vis, hm=hm)
    if flip:
        d["hmf"] = (torch.flip(hm, dims=[-1]) + 0.02 * torch.randn(B, K, H, W, generator=g, device=dev)).contiguous()
    if offsets:
        d["off"] = torch.randn(B, K, 2, H, W, generator=g, device=dev).mul_(0.3)
    if var:
        d["var"] = torch.nn.functional.softplus(torch.randn(B, K, H, W, generator=g, device=dev))
    return d


out = []
def report(name, cfg, tiles, bytes_per_tile, ms):
    gbs = tiles * bytes_per_tile / (ms * 1e-3) / 1e9
    row = dict(kernel=name, config=cfg, tiles=tiles, bytes_per_tile=bytes_per_tile, ms=round(ms, 4),
               heatmaps_per_s=round(tiles / (ms * 1e-3)), achieved_GBps=round(gbs, 1), peak_GBps=PEAK, frac=round(gbs / PEAK, 3))
    out.append(row)
    print(json.dumps(row), flush=True)


alpha = torch.tensor([0.5], device=dev); fw = torch.tensor([0.6224593312018546], device=dev)
perm = torch.tensor([0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15], dtype=torch.int32, device=dev)
DF = N.DECODE_REFINE | N.DECODE_APPLY_OFFSET

# configs[1]: 64x48, B=1024
B, K, H, W = 1024, 17, 64, 48
d = inputs(B, K, H, W, 192, 256, 2.0, var=True)
n = H * W
report("encode_kernel", "cfg1 64x48 B=1024", B * K, 4 * n, timeit(lambda: ops.encode(d["kps"], d["vis"], H, W, 192.0, 256.0, 2.0)))
report("decode_kernel", "cfg1 64x48 B=1024 refine+offset", B * K, 4 * n, timeit(lambda: ops.decode(d["hm"], None, None, d["off"], alpha, fw, 2, DF)))
report("argmax_kernel quarter", "cfg1 64x48 B=1024", B * K, 4 * n, timeit(lambda: ops.decode_argmax(d["hm"], N.ARGMAX_QUARTER)))
tgt, wgt = ops.encode(d["kps"], d["vis"], H, W, 192.0, 256.0, 2.0)
loss = lambda target, grads, dec: ops.fusion_loss(d["hm"], d["off"], d["var"], target, d["vis"], d["kps"], None, None, 192.0, 256.0, LAM, 2.0, 2.0,
                                                  True, SK, grads, dec, alpha, fw, 2, DF)
report("loss step (fused, on-the-fly target)", "cfg1 64x48 B=1024", B * K, 24 * n, timeit(lambda: loss(None, True, True)))
report("loss fwd+bwd (target from HBM)", "cfg1 64x48 B=1024", B * K, 28 * n, timeit(lambda: loss(tgt, True, False)))
report("loss fwd only (on-the-fly target)", "cfg1 64x48 B=1024", B * K, 8 * n, timeit(lambda: loss(None, False, False)))
vm = d["var"].mean(dim=(2, 3))
report("loss step with per-tile variance means (fused, on-the-fly target)", "cfg1 64x48 B=1024", B * K, 16 * n,
       timeit(lambda: ops.fusion_step_vmean(d["hm"], d["off"], vm, None, d["vis"], d["kps"], None, None, 192.0, 256.0, LAM, 2.0, 2.0, True, SK,
                                            True, True, alpha, fw, 2, DF)))
# float16 maps (autocast): the fused pass with the expected loss scale, the backward that finds its expectation met
# (returns inside the kernel) and the one that does not (computes the gradients again)
h16 = {k: d[k].half() for k in ("hm", "off", "var")}
scale = torch.tensor([65536.0], device=dev)
g7 = torch.zeros(7, device=dev); g7[6] = 65536.0
g7b = torch.zeros(7, device=dev); g7b[6] = 32768.0
f16 = lambda: ops.fusion_loss_f16(h16["hm"], h16["off"], h16["var"], None, d["vis"], d["kps"], None, scale, 192.0, 256.0, LAM, 2.0, 2.0, True, SK,
                                  True, True, alpha, fw, 2, DF)
res16 = f16()
b16 = lambda g: ops.fusion_loss_backward_f16(g, res16[3], res16[4], res16[5], True, h16["hm"], h16["off"], h16["var"], None, d["vis"], d["kps"], None,
                                             scale, 192.0, 256.0, LAM, 2.0, 2.0, True, SK)
report("loss step f16 (fused: losses + decode + half gradients at the expected loss scale)", "cfg1 64x48 B=1024", B * K, 12 * n, timeit(f16))
report("loss f16 backward, expectation met (no work)", "cfg1 64x48 B=1024", B * K, 0, timeit(lambda: b16(g7)))
report("loss f16 backward, loss scale changed (gradients again)", "cfg1 64x48 B=1024", B * K, 12 * n, timeit(lambda: b16(g7b)))
del h16, res16
# second-generation family on the same shapes
pred = d["hm"].abs().add_(0.01)
wts = [1.0, 0.15, 0.6]
comb = lambda grads: ops.combined_loss(pred, tgt, wgt, d["kps"], None, d["kps"], None, 0, N.CRIT_MSE, N.CRIT_SMOOTHL1, True, 1.0, 1.2, 0.5, wts, True, grads)
report("genb_tile_kernel CombinedLoss fwd+bwd", "cfg1 64x48 B=1024", B * K, 12 * n, timeit(lambda: comb(True)))
report("genb_tile_kernel CombinedLoss fwd", "cfg1 64x48 B=1024", B * K, 8 * n, timeit(lambda: comb(False)))
pred16 = pred.half()
comb16 = lambda grads: ops.combined_loss(pred16, tgt, wgt, d["kps"], None, d["kps"], scale, 0, N.CRIT_MSE, N.CRIT_SMOOTHL1, True, 1.0, 1.2, 0.5, wts, True, grads)
report("genb_tile_kernel<HALF> CombinedLoss fwd+bwd, float16 predictions / gradients (2N + 4N read, 2N written)", "cfg1 64x48 B=1024", B * K, 8 * n, timeit(lambda: comb16(True)))
report("CombinedLoss float16 the up-cast way (pred.float() -> float32 kernel -> grad.half()): what the float16 entry point replaces", "cfg1 64x48 B=1024", B * K, 8 * n,
       timeit(lambda: ops.combined_loss(pred16.float(), tgt, wgt, d["kps"], None, d["kps"], scale, 0, N.CRIT_MSE, N.CRIT_SMOOTHL1, True, 1.0, 1.2, 0.5, wts, True, True)[1].half()))
del pred16
cen = torch.rand(B, 2, device=dev) * 300 + 100
scl = torch.rand(B, 2, device=dev) * 200 + 150
report("postprocess_kernel (whole pipeline)", "cfg1 64x48 B=1024", B * K, 4 * n,
       timeit(lambda: ops.postprocess(pred, d["kps"], cen, scl, N.ARGMAX_TAYLOR, True, 256.0, 5, True, 0.3, True, 256.0, 256.0)))
report("heatmap_step_kernel (KeypointMSELoss fwd+bwd, on-the-fly target, arg-max)", "cfg1 64x48 B=1024", B * K, 8 * n,
       timeit(lambda: ops.heatmap_step(pred, None, d["vis"], d["kps"], 192.0, 256.0, 2.0, True, 0, None, True, True, N.ARGMAX_QUARTER)))
report("encode_genb_kernel clipped", "cfg1 64x48 B=1024", B * K, 4 * n, timeit(lambda: ops.encode_mode(d["kps"], d["vis"], H, W, 192.0, 256.0, 2.0, N.ENCODE_PATCH_CLIPPED)))
report("encode_genb_kernel dense", "cfg1 64x48 B=1024", B * K, 4 * n, timeit(lambda: ops.encode_mode(d["kps"], d["vis"], H, W, 192.0, 256.0, 2.0, N.ENCODE_DENSE)))
del d, tgt, wgt, pred
torch.cuda.empty_cache()

# configs[2]: 96x72 decode + flip + offsets, B=4096
B, K, H, W = (1024 if quick else 4096), 17, 96, 72
d = inputs(B, K, H, W, 288, 384, 2.0, flip=True)
n = H * W
report("decode_kernel flip", f"cfg2 96x72 B={B} flip+refine+offset", B * K, 8 * n, timeit(lambda: ops.decode(d["hm"], d["hmf"], perm, d["off"], alpha, fw, 2, DF), 10))
report("decode_kernel", f"cfg2 96x72 B={B} refine+offset", B * K, 4 * n, timeit(lambda: ops.decode(d["hm"], None, None, d["off"], alpha, fw, 2, DF), 10))
report("argmax_kernel quarter", f"cfg2 96x72 B={B}", B * K, 4 * n, timeit(lambda: ops.decode_argmax(d["hm"], N.ARGMAX_QUARTER), 10))
del d
torch.cuda.empty_cache()

# configs[3]: preemie 128x128, K=13, sigma 1.5, six-term step, per-GPU B=1024
B, K, H, W = (256 if quick else 1024), 13, 128, 128
d = inputs(B, K, H, W, 256, 256, 1.5, var=True)
n = H * W
report("loss step (fused, on-the-fly target)", f"cfg3 128x128 K=13 B={B}", B * K, 24 * n,
       timeit(lambda: ops.fusion_loss(d["hm"], d["off"], d["var"], None, d["vis"], d["kps"], None, None, 256.0, 256.0, LAM, 1.5, 1.5, True, SK,
                                      True, True, alpha, fw, 2, DF), 10))
report("encode_kernel", f"cfg3 128x128 K=13 B={B}", B * K, 4 * n, timeit(lambda: ops.encode(d["kps"], d["vis"], H, W, 256.0, 256.0, 1.5), 10))
del d
torch.cuda.empty_cache()

# configs[4]: decode-only sweep 64x48
for B in ((256, 4096) if quick else (256, 1024, 4096, 16384, 65536)):
    K, H, W = 17, 64, 48
    g = torch.Generator(device=dev).manual_seed(B)
    hm = torch.randn(B, K, H, W, generator=g, device=dev).mul_(0.2)
    off = None
    n = H * W
    report("decode_kernel (no offsets)", f"cfg4 sweep 64x48 B={B}", B * K, 4 * n, timeit(lambda: ops.decode(hm, None, None, None, alpha, None, 2, N.DECODE_REFINE), 10))
    report("argmax_kernel quarter", f"cfg4 sweep 64x48 B={B}", B * K, 4 * n, timeit(lambda: ops.decode_argmax(hm, N.ARGMAX_QUARTER), 10))
    del hm
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernels.json"), "w"), indent=1)
