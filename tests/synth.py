"""Seeded synthetic inputs for the codec path.

Everything is drawn with ``numpy.random.default_rng(seed)`` in a fixed order, so
the same (config, seed) gives the same bytes here and on the GPU box (same
image, same numpy).  ``digest`` fingerprints the inputs; golden fixtures store
it so that a drift of the generator is reported as such and not as a parity
failure.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Dict, Tuple

import numpy as np

from oracle import heatmap_codec as oc


@dataclass(frozen=True)
class CodecConfig:
    name: str
    B: int
    K: int
    heatmap_size: Tuple[int, int]      # (W, H)
    input_size: Tuple[int, int]        # (W_in, H_in)
    sigma: float

    @property
    def W(self): return self.heatmap_size[0]
    @property
    def H(self): return self.heatmap_size[1]


CONFIGS: Dict[str, CodecConfig] = {
    # BASELINE.json configs[0] shape family at fixture size
    "w32_256x192": CodecConfig("w32_256x192", 3, 17, (48, 64), (192, 256), 2.0),
    # configs[2] shape family
    "hrformer_384x288": CodecConfig("hrformer_384x288", 2, 17, (72, 96), (288, 384), 2.0),
    # configs[3]: preemie_optimized.yaml constants fed to the six-term loss
    "preemie_256": CodecConfig("preemie_256", 1, 13, (128, 128), (256, 256), 1.5),
}

EDGE_MU = (-8.0, -7.5, -7.025, -0.3, 0.0, 1.5, 4.2, 5.99)


def flip_perm(K: int, pairs=oc.COCO_FLIP_PAIRS) -> np.ndarray:
    perm = np.arange(K, dtype=np.int32)
    for a, b in pairs:
        if a < K and b < K:
            perm[a], perm[b] = b, a
    return perm


def make_keypoints(cfg: CodecConfig, rng: np.random.Generator, B: int):
    W_in, H_in = cfg.input_size
    vis = rng.choice(np.array([0.0, 1.0, 2.0], np.float32), size=(B, cfg.K), p=[0.15, 0.25, 0.60]).astype(np.float32)
    kps = np.empty((B, cfg.K, 2), np.float32)
    kps[..., 0] = rng.uniform(-0.1 * W_in, 1.1 * W_in, size=(B, cfg.K))
    kps[..., 1] = rng.uniform(-0.1 * H_in, 1.1 * H_in, size=(B, cfg.K))
    return kps, vis


def edge_keypoints(cfg: CodecConfig):
    """Deterministic set that walks mu across every truncation / clipping case
    of the encoder on both axes."""
    sx = cfg.input_size[0] / cfg.W
    sy = cfg.input_size[1] / cfg.H
    mus_x = list(EDGE_MU) + [cfg.W - 1.0, float(cfg.W), cfg.W + 5.9, cfg.W + 6.0]
    mus_y = list(EDGE_MU) + [cfg.H - 1.0, float(cfg.H), cfg.H + 5.9, cfg.H + 6.0]
    pts = [(mx * sx, my * sy) for mx in mus_x for my in mus_y]
    n = len(pts)
    B = (n + cfg.K - 1) // cfg.K
    kps = np.full((B, cfg.K, 2), 10.0, np.float32)
    kps.reshape(-1, 2)[:n] = np.asarray(pts, np.float32)
    vis = np.full((B, cfg.K), 2.0, np.float32)
    vis.reshape(-1)[1::7] = 1.0
    return kps, vis


def make_batch(cfg: CodecConfig, seed: int = 0, B: int | None = None) -> Dict[str, np.ndarray]:
    """kps, vis, target, weight, heatmaps P, offsets, variances, flipped-pass
    heatmaps, flip permutation — all float32 / int32 numpy, NCHW."""
    B = cfg.B if B is None else B
    rng = np.random.default_rng(seed)
    K, H, W = cfg.K, cfg.H, cfg.W
    kps, vis = make_keypoints(cfg, rng, B)
    target, weight = oc.encode_targets(kps, vis, cfg.heatmap_size, cfg.input_size, cfg.sigma)

    stride = np.array(cfg.input_size, np.float32) / np.array(cfg.heatmap_size, np.float32)
    jitter = rng.normal(0.0, 1.5, size=kps.shape).astype(np.float32) * stride
    shifted, _ = oc.encode_targets(kps + jitter, np.full_like(vis, 2.0), cfg.heatmap_size, cfg.input_size, cfg.sigma)
    amp = rng.uniform(0.3, 1.2, size=(B, K, 1, 1)).astype(np.float32)
    P = amp * shifted + 0.05 * rng.standard_normal((B, K, H, W), dtype=np.float32)
    # tie-break tiles: ~2 % constant, ~1 % with the maximum duplicated later in the tile
    kind = rng.uniform(size=(B, K))
    for b, k in zip(*np.nonzero(kind < 0.02)):
        P[b, k] = np.float32(rng.uniform(-0.5, 0.5))
    for b, k in zip(*np.nonzero((kind >= 0.02) & (kind < 0.03))):
        flat = P[b, k].reshape(-1)
        i = int(flat.argmax())
        j = int(rng.integers(0, flat.size))
        flat[j] = flat[i]
    P = P.astype(np.float32)

    offsets = (0.3 * rng.standard_normal((B, K, 2, H, W), dtype=np.float32)).astype(np.float32)
    z = rng.standard_normal((B, K, H, W), dtype=np.float32)
    variances = (np.log1p(np.exp(-np.abs(z))) + np.maximum(z, 0)).astype(np.float32)   # softplus
    perm = flip_perm(K)
    P_flip = (P[:, perm, :, ::-1] + 0.02 * rng.standard_normal((B, K, H, W), dtype=np.float32)).astype(np.float32)
    return dict(kps=kps, vis=vis, target=target, weight=weight, heatmaps=P, offsets=offsets,
                variances=variances, heatmaps_flip=np.ascontiguousarray(P_flip), flip_perm=perm)


def digest(batch: Dict[str, np.ndarray]) -> str:
    h = hashlib.sha256()
    for key in sorted(batch):
        a = np.ascontiguousarray(batch[key])
        h.update(key.encode()); h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


def make_genb_extras(cfg: CodecConfig, batch: Dict[str, np.ndarray], seed: int = 0) -> Dict[str, np.ndarray]:
    """Extra seeded inputs of the second-generation ("Gen-B") family: a positive prediction map (its
    losses and the linear-weight centroid normalise by the plain tile sum), regression-branch
    coordinates in both conventions the reference accepts, crop centres / scales."""
    rng = np.random.default_rng(seed + 1000)
    B, K = batch["kps"].shape[:2]
    W, H = cfg.heatmap_size
    pred = (np.abs(batch["heatmaps"]) + np.float32(0.01)).astype(np.float32)
    gt_hm = batch["kps"] * (np.array([W, H], np.float32) / np.array(cfg.input_size, np.float32))
    reg_px = (gt_hm * (256.0 / np.array([W, H], np.float32)) + rng.normal(0, 3.0, size=(B, K, 2))).astype(np.float32)
    reg_norm = np.clip(rng.uniform(0.02, 0.98, size=(B, K, 2)), 0, 1).astype(np.float32)
    coords = (gt_hm + rng.normal(0, 1.2, size=(B, K, 2))).astype(np.float32)
    refined = (gt_hm + rng.normal(0, 0.6, size=(B, K, 2))).astype(np.float32)
    center = rng.uniform(100, 400, size=(B, 2)).astype(np.float32)
    scale = rng.uniform(120, 380, size=(B, 2)).astype(np.float32)
    return dict(pred=pred, reg_px=reg_px, reg_norm=reg_norm, coords=coords, refined=refined,
                target_coords=gt_hm.astype(np.float32), center=center, scale=scale)
