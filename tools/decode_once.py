"""One decode launch per variant on a 64x48 batch (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import infantposeestimation_gaussianbias_b200 as pkg
pkg.load()
from infantposeestimation_gaussianbias_b200 import _native as N, ops
dev = torch.device("cuda", 0)
B, K, H, W = 16384, 17, 64, 48
g = torch.Generator(device=dev).manual_seed(0)
hm = torch.randn(B, K, H, W, generator=g, device=dev).mul_(0.2)
hm[:, :, H // 3, W // 3] += 8.0
alpha = torch.tensor([0.5], device=dev)
for _ in range(4):
    ops.decode(hm, None, None, None, alpha, None, 2, N.DECODE_REFINE)
    ops.decode_argmax(hm, N.ARGMAX_QUARTER)
torch.cuda.synchronize()
