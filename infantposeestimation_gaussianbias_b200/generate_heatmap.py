"""Batched Gaussian target encoding on the device.

Host-side mirror of the slot the reference leaves empty (data/generate_heatmap.py
is a blank file) with the numerics of the live encoder,
COCOPoseDataset._generate_target (datasets/coco_dataset.py:185-250).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
from torch import Tensor

from . import ops


def generate_heatmaps(keypoints: Tensor, keypoints_visible: Tensor,
                      heatmap_size: Sequence[int] = (48, 64), input_size: Sequence[int] = (192, 256),
                      sigma: float = 2.0) -> Tuple[Tensor, Tensor]:
    """keypoints (B,K,2) in input-image pixels, keypoints_visible (B,K) in {0,1,2};
    heatmap_size / input_size are (W, H) as in the reference's DataConfig.
    Returns target (B,K,H,W) and target_weight (B,K,1), both on keypoints.device (CUDA)."""
    squeeze = keypoints.dim() == 2
    if squeeze:
        keypoints, keypoints_visible = keypoints[None], keypoints_visible[None]
    W, H = int(heatmap_size[0]), int(heatmap_size[1])
    target, weight = ops.encode(keypoints.float(), keypoints_visible.float(), H, W,
                                float(input_size[0]), float(input_size[1]), float(sigma))
    if squeeze:
        target, weight = target[0], weight[0]
    return target, weight


class HeatmapGenerator:
    """Callable with the reference dataset's attribute names (input_size,
    heatmap_size, sigma, num_keypoints) so it can stand in for
    `dataset._generate_target` on device-resident batches."""

    def __init__(self, input_size=(192, 256), heatmap_size=(48, 64), sigma: float = 2.0, num_keypoints: int = 17):
        self.input_size = tuple(int(v) for v in input_size)
        self.heatmap_size = tuple(int(v) for v in heatmap_size)
        self.sigma = float(sigma)
        self.num_keypoints = int(num_keypoints)

    def __call__(self, keypoints: Tensor, keypoints_visible: Tensor) -> Tuple[Tensor, Tensor]:
        if keypoints.shape[-2] != self.num_keypoints:
            raise ValueError(f"expected {self.num_keypoints} keypoints, got {keypoints.shape[-2]}")
        return generate_heatmaps(keypoints, keypoints_visible, self.heatmap_size, self.input_size, self.sigma)
