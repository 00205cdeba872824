"""Drop-in shim: rebind the reference's codec entry points to the gbcodec CUDA ops.

The reference files stay byte-identical; `patch_reference()` is run once (from a
launcher or sitecustomize) before `train.py` / `validate.py` / `inference.py`
build their model:

    models.fusion_head.FusionPoseLoss.forward            -> FusionPoseLoss.forward  (this package)
    models.fusion_head.HeatmapRegressionHead.decode      -> head_decode
    models.pose_estimator.PoseEstimator.inference        -> inference (flip average fused into the decode)
    models.pose_estimator.PoseEstimator.decode_heatmaps  -> decode_heatmaps
    datasets.coco_dataset.COCOPoseDataset._generate_target (optional, `encode_on_device=True`)
        -> returns an empty (K,0,0) placeholder + the exact target weights; the patched loss
           then builds the target tiles inside the kernel from gt_keypoints (no target H2D copy)

Signatures and return values are those of the reference.  The
module objects are passed in (or imported by name) so that this file never
needs the reference on its import path.
"""
from __future__ import annotations

import importlib
from types import ModuleType
from typing import Dict, Optional

import numpy as np
import torch

from . import fusion_head as fh
from . import pose_estimator as pe

_ORIGINALS: Dict[str, object] = {}


def _loss_forward(self, outputs, target_heatmaps, target_weight, gt_keypoints,
                  input_size=(192, 256), heatmap_size=(48, 64)):
    """FusionPoseLoss.forward of the reference (fusion_head.py:745), routed to one kernel pass.
    `self` is the REFERENCE module: its weights / sigma are read by attribute name (:608-635)."""
    impl = getattr(self, "_gbcodec_impl", None)
    if impl is None:
        impl = fh.FusionPoseLoss(
            heatmap_weight=self.heatmap_weight, offset_weight=self.offset_weight, peak_weight=self.peak_weight,
            variance_weight=self.variance_weight, overlap_weight=self.overlap_weight, shape_weight=self.shape_weight,
            target_sigma=getattr(self, "target_sigma", getattr(getattr(self, "gaussian_constraint", None), "target_sigma", 2.0)),
            use_target_weight=getattr(self, "use_target_weight", True))
        object.__setattr__(self, "_gbcodec_impl", impl)
    return impl(outputs, target_heatmaps, target_weight, gt_keypoints, input_size, heatmap_size)


def _encode_placeholder(self, keypoints: np.ndarray, keypoints_visible: np.ndarray):
    """COCOPoseDataset._generate_target for DataLoader workers when the tiles are built on the device:
    the weights (visibility + the off-map rule, coco_dataset.py:214-229) are 17 scalar tests and stay
    in numpy; the tile tensor is an empty placeholder that survives default_collate."""
    K = self.num_keypoints
    W, H = self.heatmap_size
    stride = np.asarray(self.input_size, dtype=np.float64) / np.asarray(self.heatmap_size, dtype=np.float64)
    weight = np.asarray(keypoints_visible, dtype=np.float32).reshape(K, 1).copy()
    radius = self.sigma * 3
    for k in range(K):
        if weight[k, 0] < 0.5:
            continue
        mu = keypoints[k].astype(np.float64) / stride
        ul = (int(mu[0] - radius), int(mu[1] - radius))
        br = (int(mu[0] + radius + 1), int(mu[1] + radius + 1))
        if ul[0] >= W or ul[1] >= H or br[0] < 0 or br[1] < 0:
            weight[k, 0] = 0.0
    return np.zeros((K, 0, 0), dtype=np.float32), weight


def _head_forward_variance_means(self, x):
    """HeatmapRegressionHead.forward (fusion_head.py:275-307) with the variance branch's tail — Softplus, then the mean
    over the tile that is all the loss reads of it (:467-478) — done by one kernel on the last convolution's raw output:
    outputs['variances'] is the (B,K) tensor of per-tile means and the patched loss takes the map-less step."""
    from . import ops
    shared_feat = self.shared_layers(x)
    heatmaps = self.heatmap_branch(shared_feat)
    offsets = self.offset_branch(shared_feat)
    B, _, H, W = offsets.shape
    offsets = offsets.view(B, self.num_keypoints, 2, H, W)
    raw = shared_feat
    for layer in list(self.variance_branch)[:-1]:           # everything but the trailing nn.Softplus
        raw = layer(raw)
    return {"heatmaps": heatmaps, "offsets": offsets, "variances": ops.softplus_mean(raw.float()),
            "fusion_weight": torch.sigmoid(self.fusion_weight)}


def patch_reference(fusion_head: Optional[ModuleType] = None, pose_estimator: Optional[ModuleType] = None,
                    coco_dataset: Optional[ModuleType] = None, encode_on_device: bool = False,
                    variance_means: bool = False) -> Dict[str, object]:
    """Rebind the reference's entry points.  Modules default to `models.fusion_head`,
    `models.pose_estimator` (and `datasets.coco_dataset` when `encode_on_device`) imported by name.
    Returns the original callables (also kept for `unpatch_reference`)."""
    fusion_head = fusion_head or importlib.import_module("models.fusion_head")
    pose_estimator = pose_estimator or importlib.import_module("models.pose_estimator")
    saved = {
        "FusionPoseLoss.forward": fusion_head.FusionPoseLoss.forward,
        "HeatmapRegressionHead.decode": fusion_head.HeatmapRegressionHead.decode,
        "PoseEstimator.inference": pose_estimator.PoseEstimator.inference,
        "PoseEstimator.decode_heatmaps": pose_estimator.PoseEstimator.__dict__["decode_heatmaps"],
    }
    fusion_head.FusionPoseLoss.forward = _loss_forward
    fusion_head.HeatmapRegressionHead.decode = lambda self, outputs, apply_offset=True: fh.head_decode(self, outputs, apply_offset)
    pose_estimator.PoseEstimator.inference = lambda self, x, flip=True, flip_pairs=None: pe.inference(self, x, flip, flip_pairs)
    pose_estimator.PoseEstimator.decode_heatmaps = staticmethod(pe.decode_heatmaps)
    if variance_means:
        # opt-in (it changes what outputs['variances'] is): the head hands the loss mean_N(softplus(.)) per tile
        if type(list(fusion_head.HeatmapRegressionHead(8).variance_branch)[-1]).__name__ != "Softplus":
            raise RuntimeError("patch_reference(variance_means=True): the variance branch does not end in Softplus")
        saved["HeatmapRegressionHead.forward"] = fusion_head.HeatmapRegressionHead.forward
        fusion_head.HeatmapRegressionHead.forward = _head_forward_variance_means
    if encode_on_device:
        coco_dataset = coco_dataset or importlib.import_module("datasets.coco_dataset")
        saved["COCOPoseDataset._generate_target"] = coco_dataset.COCOPoseDataset._generate_target
        coco_dataset.COCOPoseDataset._generate_target = _encode_placeholder
    _ORIGINALS.update({k: (v, fusion_head, pose_estimator, coco_dataset) for k, v in saved.items()})
    return saved


def unpatch_reference() -> None:
    for name, (fn, fusion_head, pose_estimator, coco_dataset) in list(_ORIGINALS.items()):
        cls, attr = name.split(".")
        mod = {"FusionPoseLoss": fusion_head, "HeatmapRegressionHead": fusion_head,
               "PoseEstimator": pose_estimator, "COCOPoseDataset": coco_dataset}[cls]
        setattr(getattr(mod, cls), attr, fn)
        del _ORIGINALS[name]


# ------------------------------------------------------------------------------------------------
# second generation ("Gen-B"): utils/postprocess.py, models/losses.py, KeypointMSELoss
# ------------------------------------------------------------------------------------------------
_GENB_ORIGINALS: Dict[str, tuple] = {}
_GENB_FUNCTIONS = ("get_max_preds", "get_max_preds_with_subpixel", "fused_decode", "coordinate_refinement",
                   "filter_low_confidence", "transform_preds", "postprocess_predictions")


def _genb_forwards():
    """forward() bodies bound onto the REFERENCE's loss classes: `self` is the reference module, its
    settings are read by the attribute names models/losses.py gives them."""
    from . import losses as L

    def fused(self, pred_heatmaps, target_heatmaps, target_weight=None):
        if self.loss_type not in L._HEAT_CRIT:
            raise ValueError(f"Unsupported loss type: {self.loss_type}")
        return L._run(pred_heatmaps, target_heatmaps, target_weight, heat_crit=L._HEAT_CRIT[self.loss_type],
                      use_target_weight=self.use_target_weight, weights=(1.0, 0.0, 0.0))[0]

    def morph(self, pred_heatmaps, target_heatmaps, target_weight=None):
        return L._run(pred_heatmaps, target_heatmaps, target_weight, lam_var=self.lambda_variance,
                      lam_mean=self.lambda_mean, weights=(0.0, 1.0, 0.0), morph=True)[1]

    def joints(self, output, target, target_weight):
        return L._run(output, target, target_weight, heat_crit=L.N.CRIT_MSE_WEIGHTED,
                      use_target_weight=self.use_target_weight, heat_scale=0.5, weights=(1.0, 0.0, 0.0))[0]

    def keypoint_mse(self, pred, target, target_weight=None):
        return L._run(pred, target, target_weight, heat_crit=L.N.CRIT_MSE_WEIGHTED,
                      use_target_weight=self.use_target_weight, weights=(1.0, 0.0, 0.0))[0]

    def offset(self, pred_coords, target_coords, target_weight=None):
        name = type(self.criterion).__name__                 # nn.SmoothL1Loss | nn.L1Loss | nn.MSELoss (losses.py:145-152)
        crit = {"SmoothL1Loss": L.N.CRIT_SMOOTHL1, "L1Loss": L.N.CRIT_L1, "MSELoss": L.N.CRIT_MSE}[name]
        return L._run(coords=pred_coords, target_coords=target_coords, weight=target_weight, coord_crit=crit,
                      weights=(0.0, 0.0, 1.0))[2]

    def combined(self, predictions, targets):
        impl = getattr(self, "_gbcodec_impl", None)
        if impl is None:
            import types as _t
            cfg = _t.SimpleNamespace(LOSS=_t.SimpleNamespace(MORPH_LAMBDA=self.morph_loss.lambda_variance,
                                                             MORPH_WEIGHT=self.w_morph, REG_WEIGHT=self.w_reg))
            impl = L.CombinedLoss(cfg)
            impl.w_heatmap = self.w_heatmap
            impl.morph_loss.lambda_mean = self.morph_loss.lambda_mean
            object.__setattr__(self, "_gbcodec_impl", impl)
        return impl(predictions, targets)

    return {"FusedPoseLoss": fused, "MorphologyShapeLoss": morph, "JointsMSELoss": joints,
            "OffsetRegressionLoss": offset, "CombinedLoss": combined}, keypoint_mse


def patch_reference_genb(postprocess: Optional[ModuleType] = None, losses: Optional[ModuleType] = None,
                         pose_estimator: Optional[ModuleType] = None) -> Dict[str, object]:
    """Rebind the second-generation entry points: the seven functions of `utils.postprocess`
    (postprocess_predictions becomes one kernel), forward() of the five classes of `models.losses`
    and of `models.pose_estimator.KeypointMSELoss`.  Modules default to an import by name; pass
    `False` to skip one."""
    from . import postprocess as pp
    saved: Dict[str, object] = {}
    if postprocess is not False:
        postprocess = postprocess or importlib.import_module("utils.postprocess")
        for fn in _GENB_FUNCTIONS:
            saved[f"postprocess.{fn}"] = getattr(postprocess, fn)
            _GENB_ORIGINALS[f"postprocess.{fn}"] = (postprocess, fn, getattr(postprocess, fn))
            setattr(postprocess, fn, getattr(pp, fn))
    forwards, keypoint_mse = _genb_forwards()
    if losses is not False:
        losses = losses or importlib.import_module("models.losses")
        for cls, fwd in forwards.items():
            saved[f"losses.{cls}.forward"] = getattr(losses, cls).forward
            _GENB_ORIGINALS[f"losses.{cls}.forward"] = (getattr(losses, cls), "forward", getattr(losses, cls).forward)
            getattr(losses, cls).forward = fwd
    if pose_estimator is not False:
        pose_estimator = pose_estimator or importlib.import_module("models.pose_estimator")
        cls = pose_estimator.KeypointMSELoss
        saved["pose_estimator.KeypointMSELoss.forward"] = cls.forward
        _GENB_ORIGINALS["pose_estimator.KeypointMSELoss.forward"] = (cls, "forward", cls.forward)
        cls.forward = keypoint_mse
    return saved


def unpatch_reference_genb() -> None:
    for name, (owner, attr, fn) in list(_GENB_ORIGINALS.items()):
        setattr(owner, attr, fn)
        del _GENB_ORIGINALS[name]
