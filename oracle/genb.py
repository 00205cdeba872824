"""CPU restatement of the reference's "Gen-B" codec family.  TEST INFRASTRUCTURE.

Same rules as ``oracle/heatmap_codec.py``: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU
legs of ``bench.py`` may import this; the product never does.

What it restates (``file:line`` relative to the upstream reference tree):

  get_max_preds                 utils/postprocess.py:10-34
  get_max_preds_with_subpixel   utils/postprocess.py:37-75
  fused_decode                  utils/postprocess.py:78-135
  coordinate_refinement         utils/postprocess.py:138-184
  filter_low_confidence         utils/postprocess.py:226-238
  transform_preds               utils/postprocess.py:270-292
  postprocess_predictions       utils/postprocess.py:296-340
  fused_pose_loss               models/losses.py:10-47      (FusedPoseLoss)
  spatial_statistics, morphology_shape_loss   models/losses.py:50-135 (MorphologyShapeLoss)
  offset_regression_loss        models/losses.py:138-171    (OffsetRegressionLoss)
  joints_mse_loss               models/losses.py:174-202    (JointsMSELoss)
  combined_loss                 models/losses.py:205-290    (CombinedLoss)
  keypoint_mse_loss             models/pose_estimator.py:102-143 (KeypointMSELoss)
  encode_patch_clipped          data/coco_dataset.py:222-287 (PreemieCocoDataset._generate_heatmaps)
  encode_dense                  data/pose_transforms.py:385-457 (GenerateTarget)
  heatmap_to_image              validate.py:102-119, inference.py:143-175 (coordinate transform)

The loops of the reference are kept as loops where their scalar semantics matter (``int()``
truncation, Python-float Taylor step); the rest is vectorised torch/numpy with the same library
calls.  Pinned against the reference's own outputs by ``tests/golden/make_golden_genb.py`` →
``tests/golden/genb_*.npz`` → ``tests/test_oracle_genb.py``; the dense encoder is additionally
pinned to the three peaks the reference's ``data/test_transforms.py:342-379`` prints.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------
# decode family (utils/postprocess.py)
# ------------------------------------------------------------------------------------------
def get_max_preds(hm: torch.Tensor):
    """postprocess.py:10-34 — first maximum of every tile; (x, y) = (idx % W, idx // W)."""
    B, K, H, W = hm.shape
    maxvals, idx = torch.max(hm.reshape(B, K, -1), dim=2)
    preds = torch.zeros((B, K, 2), dtype=torch.float32)
    preds[:, :, 0] = idx % W
    preds[:, :, 1] = idx // W
    return preds, maxvals.unsqueeze(-1)


def get_max_preds_with_subpixel(hm: torch.Tensor):
    """postprocess.py:37-75 — second-order step, strict `1 < p < size-1`, differences in the
    tensor dtype, the division in Python floats, clip to +-0.5, added back in float32."""
    B, K, H, W = hm.shape
    preds, maxvals = get_max_preds(hm)
    for b in range(B):
        for k in range(K):
            t = hm[b, k]
            px, py = int(preds[b, k, 0]), int(preds[b, k, 1])
            if 1 < px < W - 1 and 1 < py < H - 1:
                dx = (t[py, px + 1] - t[py, px - 1]).item()
                dy = (t[py + 1, px] - t[py - 1, px]).item()
                dxx = (t[py, px + 1] - 2 * t[py, px] + t[py, px - 1]).item()
                dyy = (t[py + 1, px] - 2 * t[py, px] + t[py - 1, px]).item()
                if dxx < 0:
                    preds[b, k, 0] += float(np.clip(dx / (2 * abs(dxx)), -0.5, 0.5))
                if dyy < 0:
                    preds[b, k, 1] += float(np.clip(dy / (2 * abs(dyy)), -0.5, 0.5))
    return preds, maxvals


def fused_decode(hm: torch.Tensor, regression_coords: Optional[torch.Tensor] = None, centers=None, scales=None,
                 alpha: float = 0.5, image_size: float = 256.0):
    """postprocess.py:78-135.  The hard-coded 256, the batch-wide `regression_coords.max() <= 1.0`
    branch and the adaptive blend that overrides the fixed alpha are the reference's behaviour."""
    preds, maxvals = get_max_preds_with_subpixel(hm)
    H, W = hm.shape[-2:]
    if centers is not None and scales is not None:
        preds[:, :, 0] *= image_size / W
        preds[:, :, 1] *= image_size / H
    if regression_coords is not None:
        reg = regression_coords
        if reg.max() <= 1.0:
            reg = reg * image_size
        adaptive = maxvals / (maxvals + 0.1)
        preds = adaptive * preds + (1 - adaptive) * reg
    return preds, maxvals


def coordinate_refinement(hm: torch.Tensor, coords: torch.Tensor, window_size: int = 5):
    """postprocess.py:138-184 — linear-weight centroid of the window around int(coords)."""
    B, K, H, W = hm.shape
    out = coords.clone()
    half = window_size // 2
    for b in range(B):
        for k in range(K):
            x, y = int(coords[b, k, 0].item()), int(coords[b, k, 1].item())
            x0, x1 = max(0, x - half), min(W, x + half + 1)
            y0, y1 = max(0, y - half), min(H, y + half + 1)
            local = hm[b, k, y0:y1, x0:x1]
            if local.numel() == 0:
                continue
            ys = torch.arange(y0, y1, dtype=torch.float32)
            xs = torch.arange(x0, x1, dtype=torch.float32)
            w = local / (local.sum() + 1e-8)
            out[b, k, 0] = (w.sum(dim=0) * xs).sum()
            out[b, k, 1] = (w.sum(dim=1) * ys).sum()
    return out


def filter_low_confidence(preds, maxvals, threshold: float = 0.3):
    """postprocess.py:226-238"""
    mask = (maxvals > threshold).float()
    return preds * mask, mask


def transform_preds(coords, center, scale, input_size=(256, 256)):
    """postprocess.py:270-292 — model space -> original image."""
    out = coords.clone()
    for b in range(coords.shape[0]):
        sx = scale[b, 0] / input_size[0]
        sy = scale[b, 1] / input_size[1]
        out[b, :, 0] = coords[b, :, 0] * sx + center[b, 0] - scale[b, 0] / 2
        out[b, :, 1] = coords[b, :, 1] * sy + center[b, 1] - scale[b, 1] / 2
    return out


def postprocess_predictions(hm, regression_coords=None, center=None, scale=None, alpha: float = 0.5,
                            threshold: float = 0.3, window_size: int = 5):
    """postprocess.py:296-340 — fused_decode -> coordinate_refinement -> filter -> transform."""
    preds, maxvals = fused_decode(hm, regression_coords, center, scale, alpha=alpha)
    preds = coordinate_refinement(hm, preds, window_size)
    preds, mask = filter_low_confidence(preds, maxvals, threshold)
    if center is not None and scale is not None:
        preds = transform_preds(preds, center, scale)
    return dict(preds=preds, maxvals=maxvals, mask=mask)


def heatmap_to_image(coords: np.ndarray, center: np.ndarray, scale: np.ndarray, heatmap_size: Sequence[int],
                     input_size: Sequence[int]) -> np.ndarray:
    """validate.py:31-36 and :102-119 (same arithmetic as inference.py:143-175) — heatmap px -> input px
    -> original image, numpy float32 as there:
        c *= input / heatmap;   c = c / input * scale + center - scale / 2
    heatmap_size, input_size are (W, H); coords (B,K,2), center / scale (B,2), all float32."""
    out = np.array(coords, np.float32, copy=True)
    out[:, :, 0] *= input_size[0] / heatmap_size[0]
    out[:, :, 1] *= input_size[1] / heatmap_size[1]
    center, scale = np.asarray(center, np.float32), np.asarray(scale, np.float32)
    for ax in range(2):
        out[:, :, ax] = out[:, :, ax] / input_size[ax] * scale[:, None, ax] + center[:, None, ax] - scale[:, None, ax] / 2
    return out


# ------------------------------------------------------------------------------------------
# losses (models/losses.py, models/pose_estimator.py:102-143)
# ------------------------------------------------------------------------------------------
def fused_pose_loss(pred, target, weight=None, use_target_weight=True, loss_type="mse"):
    """losses.py:10-47 — mean over B*K*H*W of criterion(p, t) * w."""
    B, K = pred.shape[:2]
    if loss_type == "mse":
        loss = F.mse_loss(pred, target, reduction="none")
    elif loss_type == "smoothl1":
        loss = F.smooth_l1_loss(pred, target, reduction="none")
    else:
        raise ValueError(f"Unsupported loss type: {loss_type}")
    if use_target_weight and weight is not None:
        loss = loss * weight.view(B, K, 1, 1)
    return loss.mean()


def spatial_statistics(hm):
    """losses.py:70-106 — q = h / (sum h + 1e-8); mean = sum q*(x,y); var = sum q*((x,y)-mean)^2."""
    B, K, H, W = hm.shape
    flat = hm.view(B, K, -1)
    prob = (flat / (flat.sum(dim=2, keepdim=True) + 1e-8)).view(B, K, H, W)
    ys = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1).to(hm.dtype)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).to(hm.dtype)
    mean_y = (prob * ys).sum(dim=[2, 3])
    mean_x = (prob * xs).sum(dim=[2, 3])
    var_y = (prob * (ys - mean_y.view(B, K, 1, 1)) ** 2).sum(dim=[2, 3])
    var_x = (prob * (xs - mean_x.view(B, K, 1, 1)) ** 2).sum(dim=[2, 3])
    return torch.stack([mean_x, mean_y], dim=2), torch.stack([var_x, var_y], dim=2)


def morphology_shape_loss(pred, target, weight=None, lambda_variance=1.0, lambda_mean=0.5):
    """losses.py:108-135"""
    pm, pv = spatial_statistics(pred)
    tm, tv = spatial_statistics(target)
    loss = lambda_variance * F.mse_loss(pv, tv, reduction="none") + lambda_mean * F.mse_loss(pm, tm, reduction="none")
    if weight is not None:
        loss = loss * weight.view(loss.shape[0], loss.shape[1], 1)
    return loss.mean()


def offset_regression_loss(pred_coords, target_coords, weight=None, loss_type="smoothl1"):
    """losses.py:138-171"""
    crit = {"smoothl1": F.smooth_l1_loss, "l1": F.l1_loss, "mse": F.mse_loss}
    if loss_type not in crit:
        raise ValueError(f"Unsupported loss type: {loss_type}")
    loss = crit[loss_type](pred_coords, target_coords, reduction="none")
    if weight is not None:
        loss = loss * weight.view(loss.shape[0], loss.shape[1], 1)
    return loss.mean()


def joints_mse_loss(output, target, weight, use_target_weight=True):
    """losses.py:174-202 — sum_k 0.5 * mean_{b,n}((p*w - t*w)^2) / K."""
    B, K = output.shape[:2]
    p = output.reshape(B, K, -1)
    t = target.reshape(B, K, -1)
    loss = 0
    for k in range(K):
        pk, tk = p[:, k], t[:, k]
        if use_target_weight:
            loss = loss + 0.5 * F.mse_loss(pk * weight[:, k], tk * weight[:, k])
        else:
            loss = loss + 0.5 * F.mse_loss(pk, tk)
    return loss / K


def keypoint_mse_loss(pred, target, weight=None, use_target_weight=True):
    """pose_estimator.py:102-143 — mean((p*w - t*w)^2) over B*K*N."""
    B, K = pred.shape[:2]
    p, t = pred.reshape(B, K, -1), target.reshape(B, K, -1)
    if use_target_weight and weight is not None:
        return F.mse_loss(p * weight, t * weight)
    return F.mse_loss(p, t)


COMBINED_KEYS = ("heatmap", "morph", "regression", "refined", "total")


def combined_loss(predictions: Dict[str, torch.Tensor], targets: Dict[str, torch.Tensor], morph_lambda=1.0,
                  morph_weight=0.1, reg_weight=0.5, w_heatmap=1.0):
    """losses.py:205-290 — CombinedLoss.forward; missing terms count as 0 in the total."""
    w = targets.get("weights")
    losses = {}
    if "heatmaps" in predictions and "heatmaps" in targets:
        losses["heatmap"] = fused_pose_loss(predictions["heatmaps"], targets["heatmaps"], w)
        losses["morph"] = morphology_shape_loss(predictions["heatmaps"], targets["heatmaps"], w, morph_lambda, 0.5)
    if "coords" in predictions and "coords" in targets:
        losses["regression"] = offset_regression_loss(predictions["coords"], targets["coords"], w)
    if "refined_coords" in predictions and "coords" in targets:
        losses["refined"] = offset_regression_loss(predictions["refined_coords"], targets["coords"], w)
    total = (w_heatmap * losses.get("heatmap", 0) + morph_weight * losses.get("morph", 0)
             + reg_weight * losses.get("regression", 0) + reg_weight * losses.get("refined", 0))
    losses["total"] = total
    return total, losses


# ------------------------------------------------------------------------------------------
# encoders
# ------------------------------------------------------------------------------------------
def encode_patch_clipped(joints: np.ndarray, joints_vis: np.ndarray, heatmap_size: Sequence[int],
                         image_size: Sequence[int], sigma: float):
    """data/coco_dataset.py:222-287 for a batch.  heatmap_size = (H, W), image_size = (W_in, H_in)
    as in that file.  joints (B,K,2) float32, joints_vis (B,K).  Weight binarised to 1.0; the joint
    must lie inside the map (:250); `ul` is clamped to 0 BEFORE the patch slice is derived (:262-263
    vs :277), so for mu < 3 sigma the patch's own top-left corner lands on pixel 0."""
    joints = np.asarray(joints, np.float32)
    B, K = joints.shape[:2]
    H, W = int(heatmap_size[0]), int(heatmap_size[1])
    target = np.zeros((B, K, H, W), np.float32)
    weight = np.zeros((B, K, 1), np.float32)
    scale_x = W / image_size[0]
    scale_y = H / image_size[1]
    tmp = sigma * 3
    size = 2 * tmp + 1
    x = np.arange(0, size, 1, np.float32)
    y = x[:, None]
    x0 = y0 = size // 2
    g = np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
    for b in range(B):
        for k in range(K):
            if not joints_vis[b, k] > 0:
                continue
            weight[b, k] = 1.0
            mu_x = joints[b, k, 0] * scale_x           # float32 * weak Python float -> float32
            mu_y = joints[b, k, 1] * scale_y
            if mu_x < 0 or mu_y < 0 or mu_x >= W or mu_y >= H:
                weight[b, k] = 0.0
                continue
            ul = [int(mu_x - tmp), int(mu_y - tmp)]
            br = [int(mu_x + tmp + 1), int(mu_y + tmp + 1)]
            ul = [max(0, ul[0]), max(0, ul[1])]
            br = [min(W, br[0]), min(H, br[1])]
            g_x = max(0, -ul[0]), min(br[0], W) - ul[0]
            g_y = max(0, -ul[1]), min(br[1], H) - ul[1]
            img_x = max(0, ul[0]), min(br[0], W)
            img_y = max(0, ul[1]), min(br[1], H)
            target[b, k, img_y[0]:img_y[1], img_x[0]:img_x[1]] = g[g_y[0]:g_y[1], g_x[0]:g_x[1]]
    return target, weight


def encode_dense(keypoints: np.ndarray, visible: np.ndarray, heatmap_size: Sequence[int],
                 input_size: Sequence[int], sigma: float):
    """data/pose_transforms.py:385-457 for a batch.  heatmap_size = (h, w), input_size = (h, w) as
    that class unpacks them (:425-426).  Sub-pixel centre, full-tile exp, weight 1/0."""
    kps = np.asarray(keypoints, np.float32).copy()
    B, K = kps.shape[:2]
    h, w = int(heatmap_size[0]), int(heatmap_size[1])
    ih, iw = input_size
    kps[..., 0] *= w / iw
    kps[..., 1] *= h / ih
    xs = np.arange(0, w, 1, dtype=np.float32)
    ys = np.arange(0, h, 1, dtype=np.float32)[:, None]
    heat = np.zeros((B, K, h, w), np.float32)
    weight = np.ones((B, K), np.float32)
    for b in range(B):
        for k in range(K):
            if visible[b, k] > 0:
                c = kps[b, k, :2]
                if 0 <= c[0] < w and 0 <= c[1] < h:
                    g = np.exp(-((xs - c[0]) ** 2 + (ys - c[1]) ** 2) / (2 * sigma ** 2))
                    heat[b, k] = np.maximum(heat[b, k], g)
                else:
                    weight[b, k] = 0.0
            else:
                weight[b, k] = 0.0
    return heat, weight
