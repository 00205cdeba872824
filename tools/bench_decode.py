#!/usr/bin/env python
"""Decode-only roofline (BASELINE configs[2] and configs[4]): decode_kernel with / without flip and
offsets, arg-max, whole Gen-B pipeline.  GPU box only.   python tools/bench_decode.py [--quick]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import infantposeestimation_gaussianbias_b200 as pkg
pkg.load()
from infantposeestimation_gaussianbias_b200 import _native as N, ops

dev = torch.device("cuda", 0)
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
quick = "--quick" in sys.argv


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def report(name, cfg, tiles, bpt, ms):
    gbs = tiles * bpt / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel=name, config=cfg, tiles=tiles, bytes_per_tile=bpt, ms=round(ms, 4), heatmaps_per_s=round(tiles / (ms * 1e-3)),
                          achieved_GBps=round(gbs, 1), frac=round(gbs / PEAK, 3))), flush=True)


alpha = torch.tensor([0.5], device=dev); fw = torch.tensor([0.6224593312018546], device=dev)
perm = torch.tensor([0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15], dtype=torch.int32, device=dev)
DF = N.DECODE_REFINE | N.DECODE_APPLY_OFFSET
for (H, W, Bs) in ((64, 48, (1024, 16384) if quick else (256, 1024, 4096, 16384, 65536)), (96, 72, (4096,)), (128, 128, (2048,))):
    for B in Bs:
        K = 17 if H != 128 else 13
        g = torch.Generator(device=dev).manual_seed(B)
        hm = torch.randn(B, K, H, W, generator=g, device=dev).mul_(0.2)
        # a peaked tile so that the soft-argmax is not pinned to the centre
        hm[:, :, H // 3, W // 3] += 8.0
        n = H * W
        cfg = f"{H}x{W} B={B}"
        report("decode (refine)", cfg, B * K, 4 * n, timeit(lambda: ops.decode(hm, None, None, None, alpha, None, 2, N.DECODE_REFINE), 10))
        report("argmax quarter", cfg, B * K, 4 * n, timeit(lambda: ops.decode_argmax(hm, N.ARGMAX_QUARTER), 10))
        if B <= 4096:
            off = torch.randn(B, K, 2, H, W, generator=g, device=dev).mul_(0.3)
            report("decode (refine+offset)", cfg, B * K, 4 * n, timeit(lambda: ops.decode(hm, None, None, off, alpha, fw, 2, DF), 10))
            if K == 17:
                hmf = torch.flip(hm, dims=[-1]).contiguous()
                report("decode (flip+refine+offset)", cfg, B * K, 8 * n, timeit(lambda: ops.decode(hm, hmf, perm, off, alpha, fw, 2, DF), 10))
                del hmf
            report("postprocess pipeline", cfg, B * K, 4 * n,
                   timeit(lambda: ops.postprocess(hm, None, None, None, N.ARGMAX_TAYLOR, False, 256.0, 5, True, 0.3, False, 256.0, 256.0), 10))
            del off
        del hm
        torch.cuda.empty_cache()
