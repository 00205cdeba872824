"""Host-side mirror of utils/postprocess.py's heatmap decoders (Gen-B family).

    get_max_preds                 utils/postprocess.py:10-34
    get_max_preds_with_subpixel   utils/postprocess.py:37-75
    coordinate_refinement         utils/postprocess.py:138-184
    fused_decode                  utils/postprocess.py:78-135
    filter_low_confidence         utils/postprocess.py:226-238
    transform_preds               utils/postprocess.py:270-292
    postprocess_predictions       utils/postprocess.py:296-340   (one kernel for the whole pipeline)
    heatmap_to_image              validate.py:31-36,102-119; inference.py:143-175
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _native as N
from . import ops
from .fusion_head import _f32


def get_max_preds(batch_heatmaps: Tensor):
    c, v, _ = ops.fast.decode_argmax(_f32(batch_heatmaps), N.ARGMAX_PLAIN)
    return c, v.unsqueeze(-1)


def get_max_preds_with_subpixel(batch_heatmaps: Tensor):
    c, v, _ = ops.fast.decode_argmax(_f32(batch_heatmaps), N.ARGMAX_TAYLOR)
    return c, v.unsqueeze(-1)


def coordinate_refinement(heatmaps: Tensor, initial_coords: Tensor, window_size: int = 5) -> Tensor:
    return ops.fast.refine_centroid(_f32(heatmaps), _f32(initial_coords), int(window_size))


def fused_decode(heatmaps: Tensor, regression_coords=None, centers=None, scales=None, alpha: float = 0.5):
    """Keeps the reference's behaviour, including its hard-coded 256 and the fact that the
    confidence-adaptive blend overrides the fixed alpha (postprocess.py:105-131).  The reference
    decides whether to rescale `regression_coords` from a host read of its maximum (:119); here
    that test runs on the device inside the same launch sequence, no sync."""
    p, v, _ = ops.fast.postprocess(_f32(heatmaps), _f32(regression_coords), None, None, N.ARGMAX_TAYLOR,
                              centers is not None and scales is not None, 256.0, 0, False, 0.0, False, 256.0, 256.0)
    return p, v.unsqueeze(-1)


def filter_low_confidence(preds: Tensor, maxvals: Tensor, threshold: float = 0.3):
    mask = (maxvals > threshold).float()
    return preds * mask, mask


def transform_preds(coords: Tensor, center: Tensor, scale: Tensor, output_size=None, input_size=(256, 256)) -> Tensor:
    """postprocess.py:270-292: coords * scale / input_size + center - scale / 2 (output_size is unused there too)."""
    s = _f32(scale)
    k = torch.tensor([float(input_size[0]), float(input_size[1])], dtype=torch.float32, device=coords.device)
    return _f32(coords) * (s / k)[:, None, :] + _f32(center)[:, None, :] - (s / 2)[:, None, :]


def postprocess_predictions(outputs, batch_meta, config):
    """postprocess.py:296-340 — fused_decode -> coordinate_refinement -> filter_low_confidence ->
    transform_preds, all inside ONE kernel: the heatmaps are read from HBM once."""
    heatmaps = outputs["heatmaps"]
    center, scale = batch_meta.get("center"), batch_meta.get("scale")
    both = center is not None and scale is not None
    dev = heatmaps.device
    to = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float32, device=dev)
    p, v, m = ops.fast.postprocess(_f32(heatmaps), _f32(outputs.get("coords", None)), to(center) if both else None,
                              to(scale) if both else None, N.ARGMAX_TAYLOR, both, 256.0, 5, True, 0.3,
                              "center" in batch_meta and "scale" in batch_meta, 256.0, 256.0)
    return {"preds": p, "maxvals": v.unsqueeze(-1), "mask": m.unsqueeze(-1)}


def heatmap_to_image(coords: Tensor, center: Tensor, scale: Tensor, heatmap_size, input_size) -> Tensor:
    """validate.py:102-119 / inference.py:143-175 on the device: heatmap px -> input px -> original
    image, in the reference's order of float32 operations.  heatmap_size, input_size are (W, H)."""
    return ops.fast.coords_to_image(_f32(coords), _f32(center), _f32(scale), int(heatmap_size[1]), int(heatmap_size[0]),
                               float(input_size[0]), float(input_size[1]))
