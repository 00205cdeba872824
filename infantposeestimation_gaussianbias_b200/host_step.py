"""Fused codec step on HOST buffers: pinned host tensors in, host results out.

The batch is cut into chunks of whole images; while chunk i is in the loss kernel,
chunk i+1 is on its way over PCIe (copy stream + double-buffered device staging).
Only the heatmaps and the variance maps are staged: of the offset maps (half of the
input bytes) the step touches 16 floats per tile, at positions that are known only once the
tile's soft-argmax is, so the kernel reads those taps straight from the pinned host
buffer over PCIe (zero-copy) instead of shipping 2*H*W floats per tile.
Chunks are shards in the sense of sharded.py: the two batch-global normalisers
are computed first from the (tiny) keypoint/visibility arrays of the whole
batch, every chunk is then normalised by them, and the per-chunk loss vectors
simply add up — so the result equals one pass over the whole batch.
Gradients stay on the device (their consumer, the backbone backward, is there): every chunk writes its slice of three
persistent (B, ...) gradient tensors (`grad_hm`, `grad_off`, `grad_var`).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native as N
from . import ops
from .fusion_head import SKELETON


class HostCodecStep:
    def __init__(self, B: int, K: int, H: int, W: int, input_size: Sequence[int] = (192, 256), sigma: float = 2.0,
                 lambdas: Sequence[float] = (1.0, 1.0, 0.5, 0.1, 0.05, 0.05), chunk_images: int = 128,
                 device: Optional[torch.device] = None, with_grads: bool = True,
                 skeleton: Sequence[Tuple[int, int]] = SKELETON, stage_offsets: bool = False):
        self.B, self.K, self.H, self.W = B, K, H, W
        self.in_w, self.in_h = float(input_size[0]), float(input_size[1])
        self.sigma = float(sigma)
        self.lambdas = [float(v) for v in lambdas]
        self.chunk = min(chunk_images, B)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.with_grads = with_grads
        self.pairs = ops.pairs_flat([(i, j) for (i, j) in skeleton if i < K and j < K])
        d, f = self.device, torch.float32
        C = self.chunk
        self.stage_offsets = stage_offsets
        self.stage = [dict(hm=torch.empty((C, K, H, W), dtype=f, device=d), var=torch.empty((C, K, H, W), dtype=f, device=d),
                           off=torch.empty((C, K, 2, H, W), dtype=f, device=d) if stage_offsets else None) for _ in range(2)]
        self.kps_d = torch.empty((B, K, 2), dtype=f, device=d)
        self.vis_d = torch.empty((B, K), dtype=f, device=d)
        self.losses_d = torch.zeros(7, dtype=f, device=d)
        self.coords_d = torch.empty((B, K, 2), dtype=f, device=d)
        self.scores_d = torch.empty((B, K), dtype=f, device=d)
        self.alpha = torch.tensor([0.5], dtype=f, device=d)
        self.fw = torch.tensor([0.6224593312018546], dtype=f, device=d)
        # gradients of the WHOLE batch, written chunk by chunk (with_grads=False: the loss / decode only, nothing stored)
        self.grad_hm = torch.empty((B, K, H, W), dtype=f, device=d) if with_grads else None
        self.grad_off = torch.empty((B, K, 2, H, W), dtype=f, device=d) if with_grads else None
        self.grad_var = torch.empty((B, K, H, W), dtype=f, device=d) if with_grads else None
        self.chunk_losses = torch.empty((2, 7), dtype=f, device=d)
        self.ws = [torch.empty(N.lib().gbcodec_loss_workspace_bytes(C, K, H, W), dtype=torch.uint8, device=d) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=d)
        self.staged = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.out_h = dict(losses=torch.empty(7, dtype=f).pin_memory(), coords=torch.empty((B, K, 2), dtype=f).pin_memory(),
                          scores=torch.empty((B, K), dtype=f).pin_memory())
        self.launches = 0

    @property
    def grads(self):
        """(grad_hm, grad_off, grad_var) of the whole batch on the device, or None with with_grads=False."""
        return (self.grad_hm, self.grad_off, self.grad_var) if self.with_grads else None

    @property
    def h2d_bytes(self) -> int:
        n = self.B * self.K * self.H * self.W * 4                  # one (B,K,H,W) fp32 tensor
        taps = self.B * self.K * 16 * 32                           # zero-copy: 16 taps per tile, a 32-byte sector each
        return n * 2 + (n * 2 if self.stage_offsets else taps) + self.B * self.K * 12

    @property
    def d2h_bytes(self) -> int:
        return 7 * 4 + self.B * self.K * 12

    def set_decode_params(self, alpha_param: Tensor, fusion_weight: Tensor):
        self.alpha.copy_(alpha_param.reshape(1)); self.fw.copy_(fusion_weight.reshape(1))

    def __call__(self, hm_h: Tensor, off_h: Tensor, var_h: Tensor, kps_h: Tensor, vis_h: Tensor) -> Dict[str, Tensor]:
        """All arguments are pinned host tensors.  Returns pinned host tensors
        {losses (7,), coords (B,K,2), scores (B,K)}; synchronises before returning."""
        B, K, H, W, C = self.B, self.K, self.H, self.W, self.chunk
        main = torch.cuda.current_stream(self.device)
        launched0 = N.lib().gbcodec_launch_count()
        self.kps_d.copy_(kps_h, non_blocking=True)
        self.vis_d.copy_(vis_h.reshape(B, K), non_blocking=True)
        den = ops.fast.loss_denominators(self.vis_d, self.kps_d, False, H, W, self.in_w, self.in_h, self.sigma, self.pairs)
        self.losses_d.zero_()
        self.copy_stream.wait_stream(main)
        nchunk = (B + C - 1) // C
        for c in range(nchunk):
            lo, hi = c * C, min(B, (c + 1) * C)
            n = hi - lo
            s = self.stage[c & 1]
            with torch.cuda.stream(self.copy_stream):
                if c >= 2:
                    self.copy_stream.wait_event(self.consumed[c & 1])
                s["hm"][:n].copy_(hm_h[lo:hi], non_blocking=True)
                if self.stage_offsets:
                    s["off"][:n].copy_(off_h[lo:hi], non_blocking=True)
                s["var"][:n].copy_(var_h[lo:hi], non_blocking=True)
                self.staged[c & 1].record(self.copy_stream)
            main.wait_event(self.staged[c & 1])
            off_c = s["off"][:n] if self.stage_offsets else off_h[lo:hi]
            grads = (self.grad_hm[lo:hi], self.grad_off[lo:hi], self.grad_var[lo:hi]) if self.with_grads else None
            ops.fusion_step_into(s["hm"][:n], off_c, s["var"][:n], self.vis_d[lo:hi], self.kps_d[lo:hi], den,
                                 self.in_w, self.in_h, self.lambdas, self.sigma, self.sigma, True, self.pairs, self.alpha, self.fw, 2,
                                 N.DECODE_REFINE | N.DECODE_APPLY_OFFSET, self.chunk_losses[c & 1], grads,
                                 self.coords_d[lo:hi], self.scores_d[lo:hi], self.ws[c & 1])
            self.consumed[c & 1].record(main)
            self.losses_d += self.chunk_losses[c & 1]
        self.out_h["losses"].copy_(self.losses_d, non_blocking=True)
        self.out_h["coords"].copy_(self.coords_d, non_blocking=True)
        self.out_h["scores"].copy_(self.scores_d, non_blocking=True)
        main.synchronize()
        self.launches = int(N.lib().gbcodec_launch_count() - launched0)      # kernels of this library, counted where they are launched
        return self.out_h


class HostDecode:
    """Keypoint decode (HeatmapRegressionHead.decode, models/fusion_head.py:309-365, with the flip-test average of
    PoseEstimator.inference, models/pose_estimator.py:303-327, in front) on HOST buffers.

    Same pipeline as HostCodecStep: chunks of whole images cross PCIe on a copy stream into double-buffered device
    staging while the previous chunk is in the decode kernel; the offset maps stay in pinned host memory (the kernel
    reads its 8 taps per tile in place); coordinates and scores come back to pinned host memory.  Nothing couples two
    chunks, so the result is that of one call over the whole batch.
    """

    def __init__(self, B: int, K: int, H: int, W: int, flip: bool = False, apply_offset: bool = True, refine: bool = True,
                 radius: int = 2, chunk_images: int = 256, device: Optional[torch.device] = None,
                 flip_pairs: Optional[Sequence[Tuple[int, int]]] = None):
        from .pose_estimator import flip_permutation
        self.B, self.K, self.H, self.W = B, K, H, W
        self.flip, self.apply_offset, self.refine, self.radius = flip, apply_offset, refine, radius
        self.chunk = min(chunk_images, B)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        d, f, C = self.device, torch.float32, self.chunk
        self.stage = [dict(hm=torch.empty((C, K, H, W), dtype=f, device=d),
                           flip=torch.empty((C, K, H, W), dtype=f, device=d) if flip else None) for _ in range(2)]
        self.perm = flip_permutation(K, flip_pairs, d) if flip else None
        self.flags = (N.DECODE_REFINE if refine else 0) | (N.DECODE_APPLY_OFFSET if apply_offset else 0)
        self.alpha = torch.tensor([0.5], dtype=f, device=d)
        self.fw = torch.tensor([0.6224593312018546], dtype=f, device=d)
        self.coords_d = torch.empty((B, K, 2), dtype=f, device=d)
        self.scores_d = torch.empty((B, K), dtype=f, device=d)
        self.copy_stream = torch.cuda.Stream(device=d)
        self.staged = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.out_h = dict(coords=torch.empty((B, K, 2), dtype=f).pin_memory(), scores=torch.empty((B, K), dtype=f).pin_memory())
        self.launches = 0

    @property
    def h2d_bytes(self) -> int:
        n = self.B * self.K * self.H * self.W * 4
        taps = self.B * self.K * 8 * 32 if self.apply_offset else 0     # zero-copy: 8 taps per tile, a 32-byte sector each
        return n * (2 if self.flip else 1) + taps

    @property
    def d2h_bytes(self) -> int:
        return self.B * self.K * 12

    def set_decode_params(self, alpha_param: Tensor, fusion_weight: Tensor):
        self.alpha.copy_(alpha_param.reshape(1)); self.fw.copy_(fusion_weight.reshape(1))

    def __call__(self, hm_h: Tensor, hm_flip_h: Optional[Tensor] = None, off_h: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """Pinned host tensors in; pinned host {coords (B,K,2), scores (B,K)} out; synchronises before returning."""
        B, C = self.B, self.chunk
        if self.flip and hm_flip_h is None:
            raise RuntimeError("gbcodec: HostDecode(flip=True) needs the heatmaps of the flipped input")
        if self.apply_offset and off_h is None:
            raise RuntimeError("gbcodec: HostDecode(apply_offset=True) needs the offset maps")
        main = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(main)
        launched0 = N.lib().gbcodec_launch_count()
        for c in range((B + C - 1) // C):
            lo, hi = c * C, min(B, (c + 1) * C)
            n = hi - lo
            s = self.stage[c & 1]
            with torch.cuda.stream(self.copy_stream):
                if c >= 2:
                    self.copy_stream.wait_event(self.consumed[c & 1])
                s["hm"][:n].copy_(hm_h[lo:hi], non_blocking=True)
                if self.flip:
                    s["flip"][:n].copy_(hm_flip_h[lo:hi], non_blocking=True)
                self.staged[c & 1].record(self.copy_stream)
            main.wait_event(self.staged[c & 1])
            coords, scores, _ = ops.fast.decode(s["hm"][:n], s["flip"][:n] if self.flip else None, self.perm,
                                           off_h[lo:hi] if self.apply_offset else None,
                                           self.alpha if self.refine else None, self.fw if self.apply_offset else None,
                                           self.radius, self.flags)
            self.consumed[c & 1].record(main)
            self.coords_d[lo:hi] = coords
            self.scores_d[lo:hi] = scores
        self.out_h["coords"].copy_(self.coords_d, non_blocking=True)
        self.out_h["scores"].copy_(self.scores_d, non_blocking=True)
        main.synchronize()
        self.launches = int(N.lib().gbcodec_launch_count() - launched0)
        return self.out_h
