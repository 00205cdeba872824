"""Closed-form forward + gradient of the six-term loss, written the way the CUDA
kernel computes it (per-tile reductions -> per-tile scalars -> one gradient
pass), in plain torch so that the algebra can be checked against autograd on the
CPU before it is trusted on the GPU.  Test helper only.

Symbols follow DESIGN.md §"Loss kernel math".
"""
from __future__ import annotations

import math

import torch

from oracle.heatmap_codec import COCO_SKELETON, DEFAULT_LAMBDAS, skeleton_for


def _tie(a, b):
    """d min(a,b)/da with ATen's even split on ties."""
    return (a < b).to(a.dtype) + 0.5 * (a == b).to(a.dtype)


def loss_closed_form(h, off, var, tgt, weight, gt_kps, input_size, lambdas=DEFAULT_LAMBDAS,
                     sigma=2.0, use_target_weight=True, skeleton=COCO_SKELETON, denominators=None):
    B, K, H, W = h.shape
    N = H * W
    dt = h.dtype
    l1, l2, l3, l4, l5, l6 = lambdas
    w = weight.reshape(B, K).to(dt)
    pairs = skeleton_for(K, skeleton)
    if denominators is None:
        D = w.sum() + 1e-8
        D5 = sum((w[:, i] * w[:, j]).sum() for i, j in pairs) + 1e-8 if pairs else torch.tensor(1e-8, dtype=dt)
    else:
        D, D5 = (torch.as_tensor(v, dtype=dt) for v in denominators)
    if use_target_weight:
        wa, Da = w, D
    else:
        wa, Da = torch.ones_like(w), torch.tensor(float(B * K), dtype=dt)

    xs = torch.arange(W, dtype=dt).view(1, 1, 1, W).expand(B, K, H, W)
    ys = torch.arange(H, dtype=dt).view(1, 1, H, 1).expand(B, K, H, W)
    S2 = lambda t: t.sum(dim=(2, 3))
    bc = lambda t: t[:, :, None, None]

    # pass 1/2: softmax moments
    m = h.amax(dim=(2, 3))
    e = torch.exp(h - bc(m))
    Z = S2(e)
    p = e / bc(Z)
    cx, cy = S2(p * xs), S2(p * ys)
    # ground truth in heatmap pixels
    gx = gt_kps[..., 0].to(dt) * (W / input_size[0])
    gy = gt_kps[..., 1].to(dt) * (H / input_size[1])

    # heatmap term
    mse = S2((h - tgt) ** 2) / N
    # offset term: bilinear read with border clamp, OOB taps read as zero (ATen)
    ccx, ccy = cx.clamp(0, W - 1), cy.clamp(0, H - 1)
    inx = ((cx >= 0) & (cx <= W - 1)).to(dt)
    iny = ((cy >= 0) & (cy <= H - 1)).to(dt)
    x0, y0 = ccx.floor().long(), ccy.floor().long()
    fx, fy = ccx - x0, ccy - y0
    x1, y1 = x0 + 1, y0 + 1
    okx, oky = (x1 < W).to(dt), (y1 < H).to(dt)
    x1c, y1c = x1.clamp(max=W - 1), y1.clamp(max=H - 1)
    O = off.reshape(B, K, 2, N)
    tap = lambda yy, xx: O.gather(3, (yy * W + xx)[:, :, None, None].expand(B, K, 2, 1))[..., 0]   # (B,K,2)
    v00, v01 = tap(y0, x0), tap(y0, x1c) * okx[..., None]
    v10, v11 = tap(y1c, x0) * oky[..., None], tap(y1c, x1c) * (okx * oky)[..., None]
    w00, w01 = (1 - fx) * (1 - fy), fx * (1 - fy)
    w10, w11 = (1 - fx) * fy, fx * fy
    samp = w00[..., None] * v00 + w01[..., None] * v01 + w10[..., None] * v10 + w11[..., None] * v11
    dsdx = ((1 - fy)[..., None] * (v01 - v00) + fy[..., None] * (v11 - v10)) * inx[..., None]
    dsdy = ((1 - fx)[..., None] * (v10 - v00) + fx[..., None] * (v11 - v01)) * iny[..., None]
    dvec = samp - torch.stack((gx - cx, gy - cy), dim=-1)
    ad = dvec.abs()
    sl1 = torch.where(ad < 1, 0.5 * dvec * dvec, ad - 0.5)
    sl1p = torch.where(ad < 1, dvec, torch.sign(dvec))
    off_t = 0.5 * (sl1[..., 0] + sl1[..., 1])
    # peak term
    peak_t = (cx - gx) ** 2 + (cy - gy) ** 2
    # variance term
    r = torch.relu(h)
    Rp = S2(r) + 1e-8
    dx, dy = xs - bc(cx), ys - bc(cy)
    ei = dx * dx + dy * dy
    v = S2(r * ei) / Rp
    s = torch.sqrt(v + 1e-8)
    mV = S2(var) / N if var is not None else None
    var_t = (s - sigma) ** 2 + ((mV - sigma) ** 2 if var is not None else 0)
    # shape term
    u = p + 1e-8
    lg = torch.log(u)
    E = -S2(p * lg)
    Estar = math.log(2 * math.pi * math.e * sigma ** 2)
    a = -lg - p / u
    PA = S2(p * a)
    shape_t = (E - Estar) ** 2
    # overlap term
    sg = torch.sigmoid(h)
    Ssum = S2(sg)
    ovl_num = torch.zeros((), dtype=dt)
    g_ovl = torch.zeros_like(h)
    for (i, j) in pairs:
        M = torch.minimum(sg[:, i], sg[:, j]).sum(dim=(1, 2))
        mm = torch.minimum(Ssum[:, i], Ssum[:, j]) + 1e-8
        rho = M / mm
        ww = w[:, i] * w[:, j]
        ovl_num = ovl_num + (torch.relu(rho - 0.5) * ww).sum()
        act = (rho > 0.5).to(dt) * l5 * ww / D5
        for a_, b_ in ((i, j), (j, i)):
            tau = _tie(sg[:, a_], sg[:, b_])
            Tau = _tie(Ssum[:, a_], Ssum[:, b_])
            g_ovl[:, a_] += (act / mm)[:, None, None] * (tau - (rho * Tau)[:, None, None]) * sg[:, a_] * (1 - sg[:, a_])

    losses = torch.stack((
        l1 * (mse * wa).sum() / Da, l2 * (off_t * wa).sum() / Da, l3 * (peak_t * wa).sum() / Da,
        l4 * (var_t * w).sum() / D, l5 * ovl_num / D5, l6 * (shape_t * w).sum() / D))
    losses = torch.cat((losses, losses.sum()[None]))

    # ---- gradients ---------------------------------------------------------
    ka = wa / Da
    kb = w / D
    A4 = l4 * kb * (s - sigma) / s
    sq = S2(r) / Rp
    dv_dcx = -2 * (S2(r * xs) / Rp - cx * sq)
    dv_dcy = -2 * (S2(r * ys) / Rp - cy * sq)
    Fx = l3 * ka * 2 * (cx - gx) + l2 * ka * 0.5 * (sl1p[..., 0] * (dsdx[..., 0] + 1) + sl1p[..., 1] * dsdx[..., 1]) + A4 * dv_dcx
    Fy = l3 * ka * 2 * (cy - gy) + l2 * ka * 0.5 * (sl1p[..., 0] * dsdy[..., 0] + sl1p[..., 1] * (dsdy[..., 1] + 1)) + A4 * dv_dcy
    g_h = (bc(l1 * ka * 2 / N) * (h - tgt)
           + bc(A4 / Rp) * (h > 0).to(dt) * (ei - bc(v))
           + bc(l6 * kb * 2 * (E - Estar)) * p * (a - bc(PA))
           + p * (dx * bc(Fx) + dy * bc(Fy))
           + g_ovl)
    g_off = torch.zeros_like(O)
    co = (l2 * ka * 0.5)[..., None] * sl1p                                   # (B,K,2)
    for (yy, xx, ww_, ok) in ((y0, x0, w00, None), (y0, x1c, w01, okx), (y1c, x0, w10, oky), (y1c, x1c, w11, okx * oky)):
        val = co * ww_[..., None] * (ok[..., None] if ok is not None else 1)
        g_off.scatter_add_(3, (yy * W + xx)[:, :, None, None].expand(B, K, 2, 1), val[..., None])
    g_var = None
    if var is not None:
        g_var = bc(l4 * kb * 2 * (mV - sigma) / N).expand(B, K, H, W).clone()
    return losses, g_h, g_off.reshape(B, K, 2, H, W), g_var
