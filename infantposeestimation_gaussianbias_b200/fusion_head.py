"""Host-side mirror of the reference's models/fusion_head.py entry points for the
codec path, routed to the gbcodec CUDA ops.

    FusionPoseLoss.forward            models/fusion_head.py:745-806
    head_decode (-> .decode)          models/fusion_head.py:309-365
    soft_argmax                       models/fusion_head.py:37-71
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _native as N
from . import ops

# models/fusion_head.py:389-394
SKELETON = ((0, 1), (0, 2), (1, 3), (2, 4), (5, 6), (5, 7), (7, 9), (6, 8), (8, 10),
            (5, 11), (6, 12), (11, 12), (11, 13), (13, 15), (12, 14), (14, 16))
LOSS_KEYS = ("heatmap_loss", "offset_loss", "peak_loss", "variance_loss", "overlap_loss", "shape_loss", "total_loss")


def _f32(t: Optional[Tensor]) -> Optional[Tensor]:
    # autocast hands the head outputs over in fp16 (train.py:171); the codec computes in fp32
    return None if t is None else (t if t.dtype == torch.float32 else t.float())


class FusionPoseLoss(nn.Module):
    """Drop-in for the reference's FusionPoseLoss: same constructor keywords
    (fusion_head.py:608-618), same forward signature and the same seven-key dict of
    lambda-weighted 0-dim tensors, differentiable w.r.t. outputs['heatmaps'|
    'offsets'|'variances'].  One kernel pass computes the losses and, when a
    gradient is needed, d(total_loss) for all three tensors.

    Extras (all optional, defaults reproduce the reference):
      target_heatmaps=None or an empty (B,K,0,0) placeholder -> the target tiles are
          generated inside the kernel from gt_keypoints + target_weight (sigma =
          `encode_sigma`, default target_sigma) and never touch HBM;
      denominators -> (2,) tensor with the global batch sums for a rank holding a shard;
      grad_scale   -> scalar tensor, the upstream gradient the caller is going to use
          (e.g. the GradScaler scale); the in-pass gradients are pre-multiplied by it
          so that backward() has nothing left to do;
      peer         -> sharded.PeerExchange of a batch-sharded job: normalisers and losses are exchanged
          with the other ranks inside the kernels (NVLink peer memory); the returned losses are global.
          With `denominators` = the global sums prefetched by `peer.denominators(...)` nothing is exchanged in front of
          the tile kernel, and with `defer_losses=True` the step only publishes its terms: the returned losses are this
          rank's share, `peer.collect_losses()` adds the shares up when the numbers are wanted.
    """

    def __init__(self, heatmap_weight: float = 1.0, offset_weight: float = 1.0, peak_weight: float = 0.5,
                 variance_weight: float = 0.1, overlap_weight: float = 0.05, shape_weight: float = 0.05,
                 target_sigma: float = 2.0, use_target_weight: bool = True,
                 encode_sigma: Optional[float] = None, skeleton: Sequence[Tuple[int, int]] = SKELETON):
        super().__init__()
        self.heatmap_weight = heatmap_weight
        self.offset_weight = offset_weight
        self.peak_weight = peak_weight
        self.variance_weight = variance_weight
        self.overlap_weight = overlap_weight
        self.shape_weight = shape_weight
        self.target_sigma = target_sigma
        self.use_target_weight = use_target_weight
        self.encode_sigma = encode_sigma
        self.skeleton = tuple((int(a), int(b)) for a, b in skeleton)

    @property
    def lambdas(self):
        return [float(self.heatmap_weight), float(self.offset_weight), float(self.peak_weight),
                float(self.variance_weight), float(self.overlap_weight), float(self.shape_weight)]

    def pairs_for(self, K: int):
        return [(i, j) for (i, j) in self.skeleton if i < K and j < K]

    def _pairs_flat(self, K: int):
        # the flattened limb pairs of a K-channel head, built once per (skeleton, K) instead of once per call
        key = (self.skeleton, K)
        cache = self.__dict__.setdefault("_pairs_flat_cache", {})
        if key not in cache:
            cache.clear()
            cache[key] = ops.pairs_flat(self.pairs_for(K))
        return cache[key]

    def forward(self, outputs: Dict[str, Tensor], target_heatmaps: Optional[Tensor], target_weight: Tensor,
                gt_keypoints: Tensor, input_size: Tuple[int, int] = (192, 256),
                heatmap_size: Tuple[int, int] = (48, 64), *, denominators: Optional[Tensor] = None,
                grad_scale: Optional[Tensor] = None, decode: Optional[dict] = None, peer=None,
                defer_losses: bool = False) -> Dict[str, Tensor]:
        # heatmap_size is accepted and ignored, as in the reference (it uses heatmaps.shape, :771)
        if self._half_maps(outputs) and denominators is None and grad_scale is None and peer is None:
            return self._forward_f16(outputs, target_heatmaps, target_weight, gt_keypoints, input_size, decode)
        v_in = outputs.get("variances")
        if v_in is not None and v_in.dim() == 2 and peer is None:
            # a head that reduces its variance branch to mean_N(V) itself (ops.softplus_mean; patch_reference(
            # variance_means=True)): the (B,K,H,W) variance map and its gradient map never exist
            hm, off = _f32(outputs["heatmaps"]), _f32(outputs["offsets"])
            K = hm.shape[1]
            if target_heatmaps is not None and target_heatmaps.numel() == 0:
                target_heatmaps = None
            with_grads = torch.is_grad_enabled() and any(t.requires_grad for t in (hm, off, v_in))
            sigma_enc = float(self.encode_sigma if self.encode_sigma is not None else self.target_sigma)
            dec = decode or {}
            res = ops.fusion_step_vmean(
                hm, off, _f32(v_in), _f32(target_heatmaps), _f32(target_weight), _f32(gt_keypoints), denominators, grad_scale,
                float(input_size[0]), float(input_size[1]), self.lambdas, float(self.target_sigma), sigma_enc, bool(self.use_target_weight),
                self._pairs_flat(K), with_grads, bool(dec), dec.get("alpha_param"), dec.get("fusion_weight"),
                int(dec.get("radius", 2)), int(dec.get("flags", N.DECODE_REFINE | N.DECODE_APPLY_OFFSET)))
            out = {k: res[0][i] for i, k in enumerate(LOSS_KEYS)}
            if dec:
                out["coords"], out["scores"] = res[4], res[5]
            return out
        hm = _f32(outputs["heatmaps"])
        off = _f32(outputs["offsets"])
        var = _f32(outputs.get("variances"))
        B, K, H, W = hm.shape
        if target_heatmaps is not None and target_heatmaps.numel() == 0:
            target_heatmaps = None
        with_grads = torch.is_grad_enabled() and any(
            t is not None and t.requires_grad for t in (hm, off, var))
        sigma_enc = float(self.encode_sigma if self.encode_sigma is not None else self.target_sigma)
        dec = decode or {}
        res = ops.fusion_loss_eager(
            hm, off, var, _f32(target_heatmaps), _f32(target_weight), _f32(gt_keypoints), denominators, grad_scale,
            float(input_size[0]), float(input_size[1]), self.lambdas, float(self.target_sigma), sigma_enc,
            bool(self.use_target_weight), self._pairs_flat(K), with_grads,
            bool(dec), dec.get("alpha_param"), dec.get("fusion_weight"), int(dec.get("radius", 2)),
            int(dec.get("flags", N.DECODE_REFINE | N.DECODE_APPLY_OFFSET)), int(peer.address) if peer is not None else 0,
            bool(defer_losses))
        losses7 = res[0]
        out = {k: losses7[i] for i, k in enumerate(LOSS_KEYS)}
        if dec:
            out["coords"], out["scores"] = res[4], res[5]
        return out


def _fusion_loss_half_methods():
    def _half_maps(self, outputs) -> bool:
        """Autocast (train.py:171): the head's three maps arrive in float16 and the tile shape has a float16 kernel."""
        hm, off, var = outputs["heatmaps"], outputs["offsets"], outputs.get("variances")
        return (hm.dtype == torch.float16 and off.dtype == torch.float16 and (var is None or var.dtype == torch.float16)
                and hm.is_cuda and tuple(hm.shape[-2:]) in ops.HALF_TILE_SHAPES)

    def _forward_f16(self, outputs, target_heatmaps, target_weight, gt_keypoints, input_size, decode):
        hm, off, var = outputs["heatmaps"], outputs["offsets"], outputs.get("variances")
        K = hm.shape[1]
        if target_heatmaps is not None and target_heatmaps.numel() == 0:
            target_heatmaps = None
        with_grads = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (hm, off, var))
        # The upstream gradient this step will most likely see is the one the last step saw (the GradScaler's scale moves
        # once in thousands of steps).  It lives on the device; the kernels compare, nothing is read back.
        state = getattr(self, "_amp_upstream", None)
        if with_grads and (state is None or state.device != hm.device):
            state = torch.ones(1, dtype=torch.float32, device=hm.device)
            object.__setattr__(self, "_amp_upstream", state)
        sigma_enc = float(self.encode_sigma if self.encode_sigma is not None else self.target_sigma)
        dec = decode or {}
        losses7, coords, scores, _, _, _, _ = ops.fusion_loss_f16_eager(
            hm, off, var, _f32(target_heatmaps), _f32(target_weight), _f32(gt_keypoints), None, state if with_grads else None,
            float(input_size[0]), float(input_size[1]), self.lambdas, float(self.target_sigma), sigma_enc,
            bool(self.use_target_weight), self._pairs_flat(K), with_grads, bool(dec), dec.get("alpha_param"),
            dec.get("fusion_weight"), int(dec.get("radius", 2)), int(dec.get("flags", N.DECODE_REFINE | N.DECODE_APPLY_OFFSET)))
        if with_grads and losses7.requires_grad:
            def remember(g, st=state):          # what total_loss actually received, kept on the device for the next step
                st.copy_(g[6:7])
            losses7.register_hook(remember)
        out = {k: losses7[i] for i, k in enumerate(LOSS_KEYS)}
        if dec:
            out["coords"], out["scores"] = coords, scores
        return out

    FusionPoseLoss._half_maps = _half_maps
    FusionPoseLoss._forward_f16 = _forward_f16


_fusion_loss_half_methods()


def soft_argmax(heatmaps: Tensor) -> Tuple[Tensor, Tensor]:
    """SoftArgmax2D.forward (beta = 1): expected pixel and raw maximum per tile."""
    c, s, _ = ops.fast.decode(_f32(heatmaps), None, None, None, None, None, 0, 0)
    return c, s


def decode_outputs(outputs: Dict[str, Tensor], alpha_param: Optional[Tensor], apply_offset: bool = True,
                   use_subpixel_refinement: bool = True, local_radius: int = 2,
                   heatmaps_of_flipped_input: Optional[Tensor] = None, flip_perm: Optional[Tensor] = None,
                   return_centre: bool = False):
    """HeatmapRegressionHead.decode on the head's output dict; optional flip-test
    average (PoseEstimator.inference) fused into the same pass."""
    hm = _f32(outputs["heatmaps"])
    flags = 0
    if use_subpixel_refinement:
        flags |= N.DECODE_REFINE
    off = fw = None
    if apply_offset:
        flags |= N.DECODE_APPLY_OFFSET
        off = _f32(outputs["offsets"])
        fw = outputs["fusion_weight"]
    coords, scores, centre = ops.fast.decode(hm, _f32(heatmaps_of_flipped_input), flip_perm, off,
                                        alpha_param if use_subpixel_refinement else None, fw, local_radius, flags)
    return (coords, scores, centre) if return_centre else (coords, scores)


def head_decode(head: nn.Module, outputs: Dict[str, Tensor], apply_offset: bool = True) -> Tuple[Tensor, Tensor]:
    """Bound as HeatmapRegressionHead.decode by patch_reference(): reads the same
    module state the reference method reads (use_subpixel_refinement,
    subpixel_refine.alpha, subpixel_refine.local_refine.local_radius)."""
    refine = bool(getattr(head, "use_subpixel_refinement", True))
    alpha = head.subpixel_refine.alpha if refine else None
    radius = head.subpixel_refine.local_refine.local_radius if refine else 2
    return decode_outputs(outputs, alpha, apply_offset, refine, radius)
