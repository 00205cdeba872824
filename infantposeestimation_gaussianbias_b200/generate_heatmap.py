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
    target, weight = ops.fast.encode(keypoints.float(), keypoints_visible.float(), H, W,
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


def generate_heatmaps_clipped(joints: Tensor, joints_vis: Tensor, heatmap_size: Sequence[int] = (64, 48),
                              image_size: Sequence[int] = (192, 256), sigma: float = 2.0) -> Tuple[Tensor, Tensor]:
    """Batched PreemieCocoDataset._generate_heatmaps of the second-generation dataset
    (data/coco_dataset.py:222-287), quirks included: heatmap_size is (H, W) and image_size (W_in, H_in)
    as in that file; weight is 1.0 for a visible joint inside the map; the patch origin is clamped
    to 0 before the patch slice is taken.  joints (B,K,2), joints_vis (B,K) or (B,K,1)."""
    B, K = joints.shape[0], joints.shape[1]
    H, W = int(heatmap_size[0]), int(heatmap_size[1])
    from . import _native as N
    return ops.fast.encode_mode(joints.float(), joints_vis.float().reshape(B, K), H, W, float(image_size[0]), float(image_size[1]),
                           float(sigma), N.ENCODE_PATCH_CLIPPED)


class GenerateTarget:
    """Batched, device-side GenerateTarget (data/pose_transforms.py:385-457): same `encoder` dict
    (input_size (h, w), heatmap_size (h, w), sigma), same result keys.  results['keypoints'] is
    (B,K,2|3) on the device, results['keypoints_visible'] (B,K) (default: all visible)."""

    def __init__(self, encoder: dict):
        self.encoder = encoder
        self.input_size = encoder.get("input_size", (256, 256))
        self.heatmap_size = encoder.get("heatmap_size", (64, 64))
        self.sigma = encoder.get("sigma", 2.0)
        self.use_udp = encoder.get("use_udp", False)

    def __call__(self, results: dict) -> dict:
        if "keypoints" not in results:
            return results
        from . import _native as N
        kps = results["keypoints"]
        squeeze = kps.dim() == 2
        if squeeze:
            kps = kps[None]
        vis = results.get("keypoints_visible")
        vis = torch.ones(kps.shape[:2], dtype=torch.float32, device=kps.device) if vis is None else vis.reshape(kps.shape[:2])
        h, w = int(self.heatmap_size[0]), int(self.heatmap_size[1])
        ih, iw = self.input_size
        heat, weight = ops.fast.encode_mode(kps[..., :2].float().contiguous(), vis.float(), h, w, float(iw), float(ih),
                                       float(self.sigma), N.ENCODE_DENSE)
        weight = weight[..., 0]
        results["heatmaps"] = heat[0] if squeeze else heat
        results["keypoint_weights"] = weight[0] if squeeze else weight
        return results
