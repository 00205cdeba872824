"""CPU restatement of the reference's heatmap-codec hot path.  TEST INFRASTRUCTURE.

Who may use this file: ``tests/``, ``__graft_entry__.smoke()`` and the
CPU-baseline / ``--impl reference`` legs of ``bench.py``, and there only as the
checker or as the thing timed on the host cores.  The product
(``infantposeestimation_gaussianbias_b200``) never imports it.

What it restates (``file:line`` relative to the upstream reference tree):

  encode_targets        datasets/coco_dataset.py:185-250   (COCOPoseDataset._generate_target)
  soft_argmax           models/fusion_head.py:37-71        (SoftArgmax2D.forward)
  local_refine(_loop)   models/fusion_head.py:84-128       (LocalGaussianRefinement.forward)
  fusion_decode         models/fusion_head.py:151-172,309-365 (SubPixelRefinement, HeatmapRegressionHead.decode)
  flip_average          models/pose_estimator.py:303-319   (PoseEstimator.inference flip branch)
  decode_heatmaps       models/pose_estimator.py:331-373   (PoseEstimator.decode_heatmaps)
  fusion_loss           models/fusion_head.py:405-559,637-806 (GaussianDistributionConstraint, FusionPoseLoss)

The arithmetic of the reference lives in third-party libraries that are not
part of the reference tree: numpy (``exp``, slicing) and PyTorch ATen
(``softmax``, ``grid_sampler_2d``, ``smooth_l1``, ``max``, ``round``).  The
reference pins only lower bounds (requirements.txt:2-6: torch>=1.10, numpy>=1.21);
the versions this restatement was pinned under are torch 2.11.0 / numpy 2.3.5
(the image's).  Where a library call *is* the algorithm (softmax,
grid_sample(bilinear, border, align_corners=True)), the same call is used here
so that the restatement inherits the library's semantics instead of guessing.

Parity pinning: the reference has no test, fixture or known-answer vector for
this path, so the restatement is pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports the reference from
``/root/reference`` in the build container, runs it on seeded inputs and commits
the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every function here against those vectors.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# models/fusion_head.py:389-394 — limb list used by the overlap term.
COCO_SKELETON: Tuple[Tuple[int, int], ...] = (
    (0, 1), (0, 2), (1, 3), (2, 4),
    (5, 6), (5, 7), (7, 9), (6, 8), (8, 10),
    (5, 11), (6, 12), (11, 12),
    (11, 13), (13, 15), (12, 14), (14, 16),
)
# configs/config.py:41-43 — left/right channel pairs swapped by the flip test.
COCO_FLIP_PAIRS: Tuple[Tuple[int, int], ...] = (
    (1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16),
)
# models/pose_estimator.py:199-208 — the weights build_model hard-codes.
DEFAULT_LAMBDAS: Tuple[float, ...] = (1.0, 1.0, 0.5, 0.1, 0.05, 0.05)
LOSS_KEYS = ("heatmap_loss", "offset_loss", "peak_loss", "variance_loss",
             "overlap_loss", "shape_loss", "total_loss")


# --------------------------------------------------------------------------
# encode
# --------------------------------------------------------------------------
def encode_targets(keypoints, visible, heatmap_size: Sequence[int],
                   input_size: Sequence[int], sigma: float = 2.0):
    """Gaussian target tiles + weights (datasets/coco_dataset.py:185-250).

    keypoints (..., K, 2) input-image pixels, visible (..., K) in {0,1,2};
    heatmap_size / input_size are (W, H).  Returns target (..., K, H, W) f32 and
    weight (..., K, 1) f32.  Quirks kept: weight is the raw visibility (:214);
    corner indices truncate toward zero (:224-225); a patch that is wholly
    off-map zeroes the weight (:227-229) but ``br == 0`` pastes nothing and
    keeps the weight; the patch has ceil(6*sigma+1) taps centred on
    floor((6*sigma+1)/2) (:232-237); mu is float64 (:208,220-221).
    """
    kps = np.asarray(keypoints, dtype=np.float32)
    vis = np.asarray(visible, dtype=np.float32)
    lead = kps.shape[:-2]
    K = kps.shape[-2]
    kps2 = kps.reshape(-1, K, 2)
    vis2 = vis.reshape(-1, K)
    n = kps2.shape[0]
    Wm, Hm = int(heatmap_size[0]), int(heatmap_size[1])
    stride = np.asarray(input_size, dtype=np.int64) / np.asarray(heatmap_size, dtype=np.int64)  # float64

    radius = sigma * 3
    extent = 2 * radius + 1
    taps = np.arange(0, extent, 1, np.float32)
    centre = extent // 2
    patch = np.exp(-((taps[None, :] - centre) ** 2 + (taps[:, None] - centre) ** 2) / (2 * sigma ** 2))
    patch = patch.astype(np.float32, copy=False)

    target = np.zeros((n, K, Hm, Wm), dtype=np.float32)
    weight = vis2.astype(np.float32).copy()

    mu = kps2.astype(np.float64) / stride  # exact widening of the f32 coords
    lo = np.trunc(mu - radius).astype(np.int64)        # int() truncation
    hi = np.trunc(mu + radius + 1).astype(np.int64)
    for b in range(n):
        for k in range(K):
            if weight[b, k] < 0.5:
                continue
            lx, ly = int(lo[b, k, 0]), int(lo[b, k, 1])
            hx, hy = int(hi[b, k, 0]), int(hi[b, k, 1])
            if lx >= Wm or ly >= Hm or hx < 0 or hy < 0:
                weight[b, k] = 0.0
                continue
            x_from, x_to = max(0, lx), min(hx, Wm)
            y_from, y_to = max(0, ly), min(hy, Hm)
            if x_to <= x_from or y_to <= y_from:
                continue  # numpy's empty-slice assignment: nothing pasted
            target[b, k, y_from:y_to, x_from:x_to] = patch[y_from - ly:y_to - ly, x_from - lx:x_to - lx]
    return target.reshape(*lead, K, Hm, Wm), weight.reshape(*lead, K, 1)


# --------------------------------------------------------------------------
# decode
# --------------------------------------------------------------------------
def _pixel_axes(H: int, W: int, dtype, device):
    xs = torch.arange(W, dtype=dtype, device=device).view(1, 1, 1, W)
    ys = torch.arange(H, dtype=dtype, device=device).view(1, 1, H, 1)
    return xs, ys


def soft_argmax(heatmaps: torch.Tensor, beta: float = 1.0):
    """Expected pixel under softmax(beta*h) and the raw tile maximum
    (models/fusion_head.py:37-71)."""
    B, K, H, W = heatmaps.shape
    prob = torch.softmax((heatmaps * beta).reshape(B, K, H * W), dim=-1).reshape(B, K, H, W)
    xs, ys = _pixel_axes(H, W, heatmaps.dtype, heatmaps.device)
    cx = (prob * xs).sum(dim=(2, 3))
    cy = (prob * ys).sum(dim=(2, 3))
    score = heatmaps.reshape(B, K, H * W).max(dim=-1).values
    return torch.stack((cx, cy), dim=-1), score


def window_centres(coords: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Integer window centre = clamp(round-half-even(c)) (fusion_head.py:104-105).
    Returned as int64 (B,K,2)."""
    px = coords[..., 0].round().clamp(0, W - 1)
    py = coords[..., 1].round().clamp(0, H - 1)
    return torch.stack((px, py), dim=-1).long()


def local_refine_loop(heatmaps: torch.Tensor, coarse: torch.Tensor, radius: int = 2) -> torch.Tensor:
    """Per-tile softmax centroid of the (2r+1)^2 window clipped to the map, one
    tile at a time like the reference does (fusion_head.py:102-126).  Slow;
    used on small cases and for the CPU baseline timing."""
    B, K, H, W = heatmaps.shape
    out = coarse.clone()
    centre = window_centres(coarse, H, W)
    for b in range(B):
        for k in range(K):
            px, py = int(centre[b, k, 0]), int(centre[b, k, 1])
            x_from, x_to = max(0, px - radius), min(W, px + radius + 1)
            y_from, y_to = max(0, py - radius), min(H, py + radius + 1)
            if x_to <= x_from or y_to <= y_from:
                continue
            win = heatmaps[b, k, y_from:y_to, x_from:x_to]
            om = torch.softmax(win.reshape(-1), dim=0).reshape(win.shape)
            xs = torch.arange(x_from, x_to, dtype=heatmaps.dtype)
            ys = torch.arange(y_from, y_to, dtype=heatmaps.dtype)
            out[b, k, 0] = (om * xs[None, :]).sum()
            out[b, k, 1] = (om * ys[:, None]).sum()
    return out


def local_refine(heatmaps: torch.Tensor, coarse: torch.Tensor, radius: int = 2,
                 centre: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched form of :func:`local_refine_loop` (same arithmetic per tile, the
    out-of-map taps carry weight exp(-inf)=0).  ``centre`` (B,K,2) int64 replaces
    the rounded soft-argmax as the window centre: tests use it to check tiles whose
    soft-argmax lies on a .5 rounding boundary under either rounding."""
    B, K, H, W = heatmaps.shape
    if centre is None:
        centre = window_centres(coarse, H, W)
    centre = centre.long()
    d = torch.arange(-radius, radius + 1)
    xs = centre[..., 0:1] + d            # (B,K,S)
    ys = centre[..., 1:2] + d
    okx = (xs >= 0) & (xs < W)
    oky = (ys >= 0) & (ys < H)
    flat = ys.clamp(0, H - 1)[..., :, None] * W + xs.clamp(0, W - 1)[..., None, :]   # (B,K,S,S)
    S = d.numel()
    win = heatmaps.reshape(B, K, H * W).gather(2, flat.reshape(B, K, S * S)).reshape(B, K, S, S)
    ok = oky[..., :, None] & okx[..., None, :]
    win = torch.where(ok, win, torch.full_like(win, float("-inf")))
    om = torch.softmax(win.reshape(B, K, S * S), dim=-1).reshape(B, K, S, S)
    lx = (om * xs[..., None, :].to(heatmaps.dtype)).sum(dim=(2, 3))
    ly = (om * ys[..., :, None].to(heatmaps.dtype)).sum(dim=(2, 3))
    return torch.stack((lx, ly), dim=-1)


def sample_offsets(offsets: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """Bilinear read of the (B,K,2,H,W) offset maps at ``coords`` with the border
    clamp, through the same ATen call the reference makes
    (fusion_head.py:344-359, 690-701)."""
    B, K, _, H, W = offsets.shape
    grid = torch.stack((2 * coords[..., 0] / (W - 1) - 1,
                        2 * coords[..., 1] / (H - 1) - 1), dim=-1)
    got = F.grid_sample(offsets.reshape(B * K, 2, H, W), grid.reshape(B * K, 1, 1, 2),
                        mode="bilinear", padding_mode="border", align_corners=True)
    return got.reshape(B, K, 2)


def flip_average(heatmaps: torch.Tensor, heatmaps_of_flipped_input: torch.Tensor,
                 flip_pairs: Optional[Sequence[Tuple[int, int]]] = COCO_FLIP_PAIRS) -> torch.Tensor:
    """Mirror the second pass back along W, swap the paired channels and average
    (pose_estimator.py:305-319).  No one-pixel shift."""
    back = torch.flip(heatmaps_of_flipped_input, dims=[-1])
    K = heatmaps.shape[1]
    perm = list(range(K))
    for a, b in (flip_pairs or ()):
        perm[a], perm[b] = b, a
    return (heatmaps + back[:, perm]) / 2


def fusion_decode(heatmaps: torch.Tensor, offsets: Optional[torch.Tensor],
                  alpha_param: float | torch.Tensor = 0.5,
                  fusion_weight: float | torch.Tensor = 0.6224593312018546,
                  apply_offset: bool = True, refine: bool = True, radius: int = 2,
                  heatmaps_of_flipped_input: Optional[torch.Tensor] = None,
                  flip_pairs: Optional[Sequence[Tuple[int, int]]] = COCO_FLIP_PAIRS,
                  loop: bool = False, centre: Optional[torch.Tensor] = None):
    """coords (B,K,2) in heatmap pixels and scores (B,K) as
    HeatmapRegressionHead.decode returns them (fusion_head.py:309-365), with the
    optional flip-test average in front (pose_estimator.py:303-327).

    ``alpha_param`` is the raw learnable scalar (sigmoid applied here, :169);
    ``fusion_weight`` is the value the head publishes, i.e. already sigmoid-ed (:306).
    """
    if heatmaps_of_flipped_input is not None:
        heatmaps = flip_average(heatmaps, heatmaps_of_flipped_input, flip_pairs)
    coords, scores = soft_argmax(heatmaps)
    if refine:
        if centre is not None:
            local = local_refine(heatmaps, coords, radius, centre=centre)
        else:
            local = (local_refine_loop if loop else local_refine)(heatmaps, coords, radius)
        a = torch.sigmoid(torch.as_tensor(alpha_param, dtype=heatmaps.dtype))
        coords = a * coords + (1 - a) * local
    if apply_offset:
        fw = torch.as_tensor(fusion_weight, dtype=heatmaps.dtype)
        coords = coords + fw * sample_offsets(offsets, coords)
    return coords, scores


def decode_heatmaps(heatmaps: torch.Tensor, shift: bool = True):
    """First-maximum pixel plus the quarter-pixel nudge toward the larger
    neighbour (pose_estimator.py:331-373).  Returns coords (B,K,2) f32,
    maxima (B,K) and the flat index (B,K) int64."""
    B, K, H, W = heatmaps.shape
    flat = heatmaps.reshape(B, K, H * W)
    idx = torch.from_numpy(np.argmax(flat.numpy(), axis=-1))      # first occurrence
    vals = flat.gather(2, idx[..., None])[..., 0]
    x = idx % W
    y = idx // W
    coords = torch.stack((x, y), dim=-1).to(torch.float32)
    if shift:
        inner = (x > 0) & (x < W - 1) & (y > 0) & (y < H - 1)
        xc, yc = x.clamp(1, W - 2), y.clamp(1, H - 2)
        at = lambda yy, xx: flat.gather(2, (yy * W + xx)[..., None])[..., 0]
        dx = torch.sign(at(yc, xc + 1) - at(yc, xc - 1))
        dy = torch.sign(at(yc + 1, xc) - at(yc - 1, xc))
        coords[..., 0] += torch.where(inner, dx * 0.25, torch.zeros_like(dx)).to(torch.float32)
        coords[..., 1] += torch.where(inner, dy * 0.25, torch.zeros_like(dy)).to(torch.float32)
    return coords, vals, idx


# --------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------
def skeleton_for(K: int, skeleton=COCO_SKELETON):
    """Limb pairs that exist for K channels (fusion_head.py:503-505)."""
    return tuple((i, j) for (i, j) in skeleton if i < K and j < K)


def loss_denominators(weight: torch.Tensor, K: int, skeleton=COCO_SKELETON):
    """The two batch-global normalisers: sum(w)+1e-8 (fusion_head.py:480,557,653,
    708,739) and sum over limbs of w_i*w_j, +1e-8 (:523-527)."""
    w = weight.reshape(weight.shape[0], K)
    pairs = skeleton_for(K, skeleton)
    d_w = w.sum() + 1e-8
    if pairs:
        I = torch.tensor([p[0] for p in pairs])
        J = torch.tensor([p[1] for p in pairs])
        d_pair = (w[:, I] * w[:, J]).sum() + 1e-8
    else:
        d_pair = torch.as_tensor(1e-8, dtype=w.dtype)
    return d_w, d_pair


def fusion_loss(heatmaps: torch.Tensor, offsets: torch.Tensor, variances: Optional[torch.Tensor],
                target: torch.Tensor, weight: torch.Tensor, gt_keypoints: torch.Tensor,
                input_size: Sequence[int] = (192, 256),
                lambdas: Sequence[float] = DEFAULT_LAMBDAS, target_sigma: float = 2.0,
                use_target_weight: bool = True, skeleton=COCO_SKELETON,
                denominators: Optional[Tuple[float, float]] = None) -> Dict[str, torch.Tensor]:
    """The seven lambda-weighted scalars of FusionPoseLoss.forward
    (fusion_head.py:745-806).  Differentiable through torch autograd exactly as
    the reference is (the soft-argmax coordinates are *not* detached, :687).

    ``denominators`` = (sum w + 1e-8, sum w_i w_j + 1e-8) replaces the batch
    normalisers; a rank of a batch-sharded job passes the global values.
    """
    B, K, H, W = heatmaps.shape
    dt = heatmaps.dtype
    w = weight.reshape(B, K).to(dt)
    if denominators is None:
        d_w, d_pair = loss_denominators(w, K, skeleton)
    else:
        d_w, d_pair = (torch.as_tensor(v, dtype=dt) for v in denominators)

    def weighted(per_tile):           # :651-655 and its copies
        if use_target_weight:
            return (per_tile * w).sum() / d_w
        return per_tile.mean()

    coords, _ = soft_argmax(heatmaps)

    # 1. heatmap MSE (:637-657)
    l_hm = weighted(((heatmaps - target) ** 2).mean(dim=(2, 3)))

    # ground truth in heatmap pixels (:679-684, 727-732); input_size is (W_in, H_in)
    gt = gt_keypoints.to(dt).clone()
    gt[..., 0] = gt_keypoints[..., 0] * (W / input_size[0])
    gt[..., 1] = gt_keypoints[..., 1] * (H / input_size[1])

    # 2. offset SmoothL1 at the predicted peak (:659-712)
    sampled = sample_offsets(offsets, coords)
    l_off = weighted(F.smooth_l1_loss(sampled, gt - coords, reduction="none").mean(dim=-1))

    # 3. peak distance (:714-743)
    l_peak = weighted(((coords - gt) ** 2).sum(dim=-1))

    # 4. variance alignment (:405-482) — these three always use the weights
    xs, ys = _pixel_axes(H, W, dt, heatmaps.device)
    pos = torch.relu(heatmaps)
    q = pos / (pos.sum(dim=(2, 3), keepdim=True) + 1e-8)
    vx = (q * (xs - coords[..., 0, None, None]) ** 2).sum(dim=(2, 3))
    vy = (q * (ys - coords[..., 1, None, None]) ** 2).sum(dim=(2, 3))
    spread = torch.sqrt(vx + vy + 1e-8)
    per_tile = (spread - target_sigma) ** 2
    if variances is not None:
        per_tile = per_tile + (variances.mean(dim=(2, 3)) - target_sigma) ** 2
    l_var = (per_tile * w).sum() / d_w

    # 5. limb overlap (:484-527)
    sg = torch.sigmoid(heatmaps)
    l_ovl = torch.zeros((), dtype=dt)
    for (i, j) in skeleton_for(K, skeleton):
        a, b = sg[:, i], sg[:, j]
        shared = torch.min(a, b).sum(dim=(1, 2))
        smaller = torch.min(a.sum(dim=(1, 2)), b.sum(dim=(1, 2))) + 1e-8
        l_ovl = l_ovl + (torch.relu(shared / smaller - 0.5) * w[:, i] * w[:, j]).sum()
    l_ovl = l_ovl / d_pair

    # 6. entropy shape (:529-559)
    p = torch.softmax(heatmaps.reshape(B, K, H * W), dim=-1)
    ent = -(p * torch.log(p + 1e-8)).sum(dim=-1)
    want = math.log(2 * math.pi * math.e * target_sigma ** 2)
    l_shape = ((ent - want) ** 2 * w).sum() / d_w

    terms = (l_hm, l_off, l_peak, l_var, l_ovl, l_shape)
    out = {k: lam * t for k, lam, t in zip(LOSS_KEYS, lambdas, terms)}
    out["total_loss"] = sum(out[k] for k in LOSS_KEYS[:-1])
    return out


def fusion_loss_and_grads(heatmaps, offsets, variances, target, weight, gt_keypoints, **kw):
    """Loss dict (detached) and d total_loss / d (heatmaps, offsets, variances)
    by autograd — the backward the reference's train.py:182 triggers."""
    h = heatmaps.detach().clone().requires_grad_(True)
    o = offsets.detach().clone().requires_grad_(True)
    v = variances.detach().clone().requires_grad_(True) if variances is not None else None
    losses = fusion_loss(h, o, v, target, weight, gt_keypoints, **kw)
    losses["total_loss"].backward()
    grads = {"heatmaps": h.grad, "offsets": o.grad if o.grad is not None else torch.zeros_like(o),
             "variances": v.grad if v is not None else None}
    return {k: t.detach() for k, t in losses.items()}, grads


def codec_step(keypoints, visible, heatmaps, offsets, variances, *, heatmap_size, input_size,
               sigma=2.0, lambdas=DEFAULT_LAMBDAS, alpha_param=0.5,
               fusion_weight=0.6224593312018546, loop_decode=True):
    """One pass of the whole path on the host: encode -> loss fwd+bwd -> decode.
    This is what the CPU baseline times."""
    target, weight = encode_targets(keypoints, visible, heatmap_size, input_size, sigma)
    losses, grads = fusion_loss_and_grads(
        heatmaps, offsets, variances, torch.from_numpy(target), torch.from_numpy(weight),
        torch.as_tensor(keypoints), input_size=input_size, lambdas=lambdas, target_sigma=sigma)
    with torch.no_grad():
        coords, scores = fusion_decode(heatmaps, offsets, alpha_param, fusion_weight, loop=loop_decode)
    return losses, grads, coords, scores
