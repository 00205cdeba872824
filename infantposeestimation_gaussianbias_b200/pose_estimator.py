"""Host-side mirror of models/pose_estimator.py's codec entry points.

    inference        PoseEstimator.inference        models/pose_estimator.py:275-329
    decode_heatmaps  PoseEstimator.decode_heatmaps  models/pose_estimator.py:331-373
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native as N
from . import ops
from .fusion_head import _f32, decode_outputs


def flip_permutation(K: int, flip_pairs: Optional[Sequence[Tuple[int, int]]], device) -> Tensor:
    perm = list(range(K))
    for a, b in (flip_pairs or ()):
        perm[a], perm[b] = b, a
    return torch.tensor(perm, dtype=torch.int32, device=device)


def decode_heatmaps(heatmaps: Tensor, shift: bool = True) -> Tuple[Tensor, Tensor]:
    """First-maximum pixel (+ quarter-pixel nudge): (keypoints (B,K,2), max_vals (B,K))."""
    c, v, _ = ops.fast.decode_argmax(_f32(heatmaps), N.ARGMAX_QUARTER if shift else N.ARGMAX_PLAIN)
    return c, v


def inference(model, x: Tensor, flip: bool = True, flip_pairs: Optional[list] = None) -> Tuple[Tensor, Tensor]:
    """PoseEstimator.inference with the backbone/head forwards left to the model
    and everything after them done in one kernel: the mirrored second pass is
    averaged in while the tile is loaded, so the flipped heatmaps are never
    materialised, channel-swapped or re-read, and the decode the reference
    throws away when flip is on (pose_estimator.py:297) is not computed."""
    output = model.forward(x)
    fusion = model.head_type == "fusion"
    use_flip = flip and flip_pairs is not None
    flipped = None
    perm = None
    if use_flip:
        flipped = model.forward(torch.flip(x, dims=[-1]))["heatmaps"]
        perm = flip_permutation(output["heatmaps"].shape[1], flip_pairs, x.device)
    if fusion:
        head = model.head
        refine = bool(getattr(head, "use_subpixel_refinement", True))
        alpha = head.subpixel_refine.alpha if refine else None
        radius = head.subpixel_refine.local_refine.local_radius if refine else 2
        return decode_outputs(output, alpha, True, refine, radius, flipped, perm)
    heatmaps = output["heatmaps"]
    if use_flip:
        back = torch.flip(flipped, dims=[-1])[:, perm.long()]
        heatmaps = (heatmaps + back) / 2
    return decode_heatmaps(heatmaps)


class HeatmapHeadStep(torch.nn.Module):
    """The training / validation step of the plain heatmap head (PoseEstimator with head_type='heatmap',
    models/pose_estimator.py:209-215,259-273) in ONE pass over the heatmaps: KeypointMSELoss forward and
    backward (:102-143), target tiles generated in the kernel from the batch's keypoints
    (datasets/coco_dataset.py:185-250) unless `target` is given, and decode_heatmaps (:331-373).

        loss, keypoints, max_vals = step(heatmaps, keypoints=gt_keypoints, keypoints_visible=vis)
        loss.backward()
    """

    def __init__(self, input_size=(192, 256), sigma: float = 2.0, use_target_weight: bool = True, shift: bool = True):
        super().__init__()
        self.input_size = tuple(float(v) for v in input_size)
        self.sigma = float(sigma)
        self.use_target_weight = bool(use_target_weight)
        self.shift = bool(shift)

    def forward(self, heatmaps: Tensor, target: Optional[Tensor] = None, target_weight: Optional[Tensor] = None,
                keypoints: Optional[Tensor] = None, keypoints_visible: Optional[Tensor] = None, *, decode: bool = True,
                norm_batch: int = 0, grad_scale: Optional[Tensor] = None):
        hm = _f32(heatmaps)
        with_grads = torch.is_grad_enabled() and hm.requires_grad
        if target is None:
            if keypoints is None or keypoints_visible is None:
                raise ValueError("HeatmapHeadStep: give `target` (+ `target_weight`) or `keypoints` + `keypoints_visible`")
            weight, kps = _f32(keypoints_visible), _f32(keypoints)
        else:
            weight, kps = _f32(target_weight), None
        loss, _, coords, maxvals = ops.heatmap_step(hm, _f32(target), weight, kps, self.input_size[0], self.input_size[1], self.sigma,
                                                    self.use_target_weight, int(norm_batch), grad_scale, with_grads, bool(decode),
                                                    N.ARGMAX_QUARTER if self.shift else N.ARGMAX_PLAIN)
        return (loss, coords, maxvals) if decode else loss
