"""Host-side mirror of utils/postprocess.py's heatmap decoders (Gen-B family).

    get_max_preds                 utils/postprocess.py:10-34
    get_max_preds_with_subpixel   utils/postprocess.py:37-75
    coordinate_refinement         utils/postprocess.py:138-184
    fused_decode                  utils/postprocess.py:78-135
    filter_low_confidence         utils/postprocess.py:226-238
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _native as N
from . import ops
from .fusion_head import _f32


def get_max_preds(batch_heatmaps: Tensor):
    c, v, _ = ops.decode_argmax(_f32(batch_heatmaps), N.ARGMAX_PLAIN)
    return c, v.unsqueeze(-1)


def get_max_preds_with_subpixel(batch_heatmaps: Tensor):
    c, v, _ = ops.decode_argmax(_f32(batch_heatmaps), N.ARGMAX_TAYLOR)
    return c, v.unsqueeze(-1)


def coordinate_refinement(heatmaps: Tensor, initial_coords: Tensor, window_size: int = 5) -> Tensor:
    return ops.refine_centroid(_f32(heatmaps), _f32(initial_coords), int(window_size))


def fused_decode(heatmaps: Tensor, regression_coords=None, centers=None, scales=None, alpha: float = 0.5):
    """Keeps the reference's behaviour, including its hard-coded 256 and the fact that
    the confidence-adaptive blend overrides the fixed alpha (postprocess.py:105-131).
    The reference decides whether to rescale `regression_coords` from a host read of
    its maximum (:119); here that decision is a device-side select, no sync."""
    preds, maxvals = get_max_preds_with_subpixel(heatmaps)
    H, W = heatmaps.shape[-2:]
    if centers is not None and scales is not None:
        preds = preds * torch.tensor([256 / W, 256 / H], dtype=preds.dtype, device=preds.device)
    if regression_coords is not None:
        reg = regression_coords.to(preds.dtype)
        reg = torch.where(reg.max() <= 1.0, reg * 256, reg)
        adaptive = maxvals / (maxvals + 0.1)
        preds = adaptive * preds + (1 - adaptive) * reg
    return preds, maxvals


def filter_low_confidence(preds: Tensor, maxvals: Tensor, threshold: float = 0.3):
    mask = (maxvals > threshold).float()
    return preds * mask, mask
