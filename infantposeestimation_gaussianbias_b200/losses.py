"""Host-side mirror of the reference's models/losses.py (and KeypointMSELoss of
models/pose_estimator.py:102-143), routed to the gbcodec CUDA ops: same class names,
constructor keywords, forward signatures, return values and ValueErrors.

    FusedPoseLoss          models/losses.py:10-47
    MorphologyShapeLoss    models/losses.py:50-135
    OffsetRegressionLoss   models/losses.py:138-171
    JointsMSELoss          models/losses.py:174-202
    CombinedLoss           models/losses.py:205-290
    build_loss             models/losses.py:293-296
    KeypointMSELoss        models/pose_estimator.py:102-143

Every forward is ONE pass over the heatmaps (read pred and target once, 8N bytes per tile) that
also leaves d(loss)/d(pred) behind (4N bytes) when a gradient is needed; CombinedLoss evaluates
all of its terms in that single pass instead of the reference's ~40 ATen launches.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _native as N
from . import ops
from .fusion_head import _f32

_HEAT_CRIT = {"mse": N.CRIT_MSE, "smoothl1": N.CRIT_SMOOTHL1}
_COORD_CRIT = {"smoothl1": N.CRIT_SMOOTHL1, "l1": N.CRIT_L1, "mse": N.CRIT_MSE}


def _needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def _run(pred=None, target=None, weight=None, coords=None, refined=None, target_coords=None, *, heat_crit=N.CRIT_MSE,
         coord_crit=N.CRIT_SMOOTHL1, use_target_weight=True, heat_scale=1.0, lam_var=1.0, lam_mean=0.5,
         weights=(1.0, 0.0, 0.0), morph=False, norm_batch=0, grad_scale=None):
    # float16 heatmaps (autocast) go to the kernel as they are: it up-casts them where it reads them and returns a float16
    # gradient rounded once (gbcodec_combined_loss_f16); everything else is float32
    keep_half = pred is not None and pred.dtype == torch.float16 and pred.is_cuda
    res = ops.combined_loss(pred if keep_half else _f32(pred), _f32(target), _f32(weight), _f32(coords), _f32(refined), _f32(target_coords),
                            grad_scale, int(norm_batch), int(heat_crit), int(coord_crit), bool(use_target_weight),
                            float(heat_scale), float(lam_var), float(lam_mean), [float(w) for w in weights], bool(morph),
                            _needs_grad(pred, coords, refined))
    return res[0]


class FusedPoseLoss(nn.Module):
    def __init__(self, use_target_weight=True, loss_type="mse"):
        super().__init__()
        if loss_type not in _HEAT_CRIT:
            raise ValueError(f"Unsupported loss type: {loss_type}")
        self.use_target_weight = use_target_weight
        self.loss_type = loss_type

    def forward(self, pred_heatmaps, target_heatmaps, target_weight=None):
        return _run(pred_heatmaps, target_heatmaps, target_weight, heat_crit=_HEAT_CRIT[self.loss_type],
                    use_target_weight=self.use_target_weight, weights=(1.0, 0.0, 0.0))[0]


class MorphologyShapeLoss(nn.Module):
    def __init__(self, lambda_variance=1.0, lambda_mean=0.5):
        super().__init__()
        self.lambda_variance = lambda_variance
        self.lambda_mean = lambda_mean

    def forward(self, pred_heatmaps, target_heatmaps, target_weight=None):
        return _run(pred_heatmaps, target_heatmaps, target_weight, lam_var=self.lambda_variance, lam_mean=self.lambda_mean,
                    weights=(0.0, 1.0, 0.0), morph=True)[1]


class OffsetRegressionLoss(nn.Module):
    def __init__(self, loss_type="smoothl1"):
        super().__init__()
        if loss_type not in _COORD_CRIT:
            raise ValueError(f"Unsupported loss type: {loss_type}")
        self.loss_type = loss_type

    def forward(self, pred_coords, target_coords, target_weight=None):
        return _run(coords=pred_coords, target_coords=target_coords, weight=target_weight,
                    coord_crit=_COORD_CRIT[self.loss_type], weights=(0.0, 0.0, 1.0))[2]


class JointsMSELoss(nn.Module):
    def __init__(self, use_target_weight=True):
        super().__init__()
        self.use_target_weight = use_target_weight

    def forward(self, output, target, target_weight):
        return _run(output, target, target_weight, heat_crit=N.CRIT_MSE_WEIGHTED, use_target_weight=self.use_target_weight,
                    heat_scale=0.5, weights=(1.0, 0.0, 0.0))[0]


class KeypointMSELoss(nn.Module):
    def __init__(self, use_target_weight: bool = True):
        super().__init__()
        self.use_target_weight = use_target_weight

    def forward(self, pred: Tensor, target: Tensor, target_weight: Optional[Tensor] = None) -> Tensor:
        return _run(pred, target, target_weight, heat_crit=N.CRIT_MSE_WEIGHTED, use_target_weight=self.use_target_weight,
                    weights=(1.0, 0.0, 0.0))[0]


class CombinedLoss(nn.Module):
    """Total = w_heatmap * L_heatmap + w_morph * L_morph + w_reg * (L_reg + L_refined); `config` is the
    reference's config object (config.LOSS.MORPH_LAMBDA / MORPH_WEIGHT / REG_WEIGHT, losses.py:216-231).
    `norm_batch` / `grad_scale`: see gbcodec_combined_desc (batch-sharded runs, loss scaling)."""

    def __init__(self, config):
        super().__init__()
        self.heatmap_loss = FusedPoseLoss(use_target_weight=True, loss_type="mse")
        self.morph_loss = MorphologyShapeLoss(lambda_variance=config.LOSS.MORPH_LAMBDA, lambda_mean=0.5)
        self.regression_loss = OffsetRegressionLoss(loss_type="smoothl1")
        self.w_heatmap = 1.0
        self.w_morph = config.LOSS.MORPH_WEIGHT
        self.w_reg = config.LOSS.REG_WEIGHT

    def forward(self, predictions: Dict[str, Tensor], targets: Dict[str, Tensor], *, norm_batch: int = 0,
                grad_scale: Optional[Tensor] = None):
        heat = "heatmaps" in predictions and "heatmaps" in targets
        reg = "coords" in predictions and "coords" in targets
        ref = "refined_coords" in predictions and "coords" in targets
        if not (heat or reg or ref):
            return 0, {"total": 0}
        l5 = _run(predictions["heatmaps"] if heat else None, targets["heatmaps"] if heat else None, targets.get("weights"),
                  predictions["coords"] if reg else None, predictions["refined_coords"] if ref else None,
                  targets["coords"] if (reg or ref) else None,
                  heat_crit=_HEAT_CRIT[self.heatmap_loss.loss_type], coord_crit=_COORD_CRIT[self.regression_loss.loss_type],
                  use_target_weight=self.heatmap_loss.use_target_weight, lam_var=self.morph_loss.lambda_variance,
                  lam_mean=self.morph_loss.lambda_mean, weights=(self.w_heatmap, self.w_morph, self.w_reg), morph=heat,
                  norm_batch=norm_batch, grad_scale=grad_scale)
        losses = {}
        if heat:
            losses["heatmap"], losses["morph"] = l5[0], l5[1]
        if reg:
            losses["regression"] = l5[2]
        if ref:
            losses["refined"] = l5[3]
        losses["total"] = l5[4]
        return l5[4], losses


def build_loss(config):
    return CombinedLoss(config)
