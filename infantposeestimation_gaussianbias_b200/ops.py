"""torch.library custom ops `gbcodec::*` over the C ABI of libgbcodec.so.

PyTorch is plumbing here: it owns the device buffers and the stream; every op
passes raw device pointers and `torch.cuda.current_stream()` to the library.
All ops require CUDA tensors and raise otherwise — there is no CPU path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native as N

_NS = "gbcodec"


class _on_device:
    """`with torch.cuda.device(dev)` only when `dev` is not the current device already (the context manager costs ~8 us of
    host time per call; one process per GPU never needs it)."""
    __slots__ = ("guard",)

    def __init__(self, dev):
        if dev.type != "cuda":
            raise RuntimeError("gbcodec: tensors must live on a CUDA device (this library has no CPU path)")
        cur = torch.cuda.current_device()
        self.guard = None if (dev.index is None or dev.index == cur) else torch.cuda.device(dev)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            return self.guard.__exit__(*exc)
        return False


def _ptr(t: Optional[Tensor]):
    return None if t is None else N._P(t.data_ptr())


def _stream(t: Tensor):
    return N._P(torch.cuda.current_stream(t.device).cuda_stream)


def _cuda_f32(name: str, t: Tensor, shape: Optional[Sequence[int]] = None) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"gbcodec: `{name}` must be a CUDA tensor (this library has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"gbcodec: `{name}` must be float32, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"gbcodec: `{name}` has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t.contiguous()


def _device_readable_f32(name: str, t: Tensor, shape: Sequence[int]) -> Tensor:
    """A CUDA tensor, or a PINNED host tensor the kernels read in place over PCIe (zero-copy; with
    unified addressing a cudaHostAlloc'd buffer has the same address on the device).  Only for
    inputs of which a kernel touches a few bytes per tile — the offset maps (8 taps of 2*H*W)."""
    if t.is_cuda:
        return _cuda_f32(name, t, shape)
    if not t.is_pinned():
        raise RuntimeError(f"gbcodec: `{name}` must be a CUDA tensor or a pinned host tensor (this library has no CPU path)")
    if t.dtype != torch.float32 or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise RuntimeError(f"gbcodec: pinned `{name}` must be contiguous float32 of shape {tuple(shape)}")
    return t


def _scalar(name: str, t: Optional[Tensor], like: Tensor) -> Optional[Tensor]:
    if t is None:
        return None
    if t.numel() != 1:
        raise RuntimeError(f"gbcodec: `{name}` must hold one element")
    if t.dtype is torch.float32 and t.device == like.device and not t.requires_grad:
        return t if t.dim() == 1 else t.reshape(1)             # (the common case costs 0.5 us instead of 5)
    return t.detach().to(device=like.device, dtype=torch.float32).reshape(1).contiguous()


# --------------------------------------------------------------------------- encode
@torch.library.custom_op(f"{_NS}::encode", mutates_args=())
def encode(kps: Tensor, vis: Tensor, H: int, W: int, in_w: float, in_h: float, sigma: float) -> Tuple[Tensor, Tensor]:
    B, K = kps.shape[0], kps.shape[1]
    kps = _cuda_f32("keypoints", kps, (B, K, 2))
    vis = _cuda_f32("visible", vis.reshape(B, K), (B, K))
    target = torch.empty((B, K, H, W), dtype=torch.float32, device=kps.device)
    weight = torch.empty((B, K, 1), dtype=torch.float32, device=kps.device)
    with _on_device(kps.device):
        N.check(N.lib().gbcodec_encode_f32(_ptr(kps), _ptr(vis), _ptr(target), _ptr(weight), B, K, H, W,
                                           in_w, in_h, sigma, _stream(kps)), "encode")
    return target, weight


@encode.register_fake
def _(kps, vis, H, W, in_w, in_h, sigma):
    B, K = kps.shape[0], kps.shape[1]
    return kps.new_empty((B, K, H, W)), kps.new_empty((B, K, 1))


# --------------------------------------------------------------------------- decode
@torch.library.custom_op(f"{_NS}::decode", mutates_args=())
def decode(hm: Tensor, hm_flipped: Optional[Tensor], flip_perm: Optional[Tensor], off: Optional[Tensor],
           alpha_param: Optional[Tensor], fusion_weight: Optional[Tensor], radius: int, flags: int
           ) -> Tuple[Tensor, Tensor, Tensor]:
    B, K, H, W = hm.shape
    hm = _cuda_f32("heatmaps", hm)
    if hm_flipped is not None:
        hm_flipped = _cuda_f32("heatmaps_flipped", hm_flipped, (B, K, H, W))
    if flip_perm is not None:
        flip_perm = flip_perm.to(device=hm.device, dtype=torch.int32).contiguous()
        if flip_perm.numel() != K:
            raise RuntimeError("gbcodec: flip_perm must hold K entries")
    if off is not None:
        off = _device_readable_f32("offsets", off, (B, K, 2, H, W))     # 8 taps per tile: may stay in pinned host memory
    alpha_param = _scalar("alpha", alpha_param, hm)
    fusion_weight = _scalar("fusion_weight", fusion_weight, hm)
    coords = torch.empty((B, K, 2), dtype=torch.float32, device=hm.device)
    scores = torch.empty((B, K), dtype=torch.float32, device=hm.device)
    centre = torch.empty((B, K, 2), dtype=torch.int32, device=hm.device)
    with _on_device(hm.device):
        N.check(N.lib().gbcodec_decode_f32(_ptr(hm), _ptr(hm_flipped), _ptr(flip_perm), _ptr(off), _ptr(alpha_param),
                                           _ptr(fusion_weight), B, K, H, W, radius, flags,
                                           _ptr(coords), _ptr(scores), _ptr(centre), _stream(hm)), "decode")
    return coords, scores, centre


@decode.register_fake
def _(hm, hm_flipped, flip_perm, off, alpha_param, fusion_weight, radius, flags):
    B, K = hm.shape[0], hm.shape[1]
    return hm.new_empty((B, K, 2)), hm.new_empty((B, K)), hm.new_empty((B, K, 2), dtype=torch.int32)


@torch.library.custom_op(f"{_NS}::decode_argmax", mutates_args=())
def decode_argmax(hm: Tensor, mode: int) -> Tuple[Tensor, Tensor, Tensor]:
    B, K, H, W = hm.shape
    hm = _cuda_f32("heatmaps", hm)
    coords = torch.empty((B, K, 2), dtype=torch.float32, device=hm.device)
    maxvals = torch.empty((B, K), dtype=torch.float32, device=hm.device)
    index = torch.empty((B, K), dtype=torch.int32, device=hm.device)
    with _on_device(hm.device):
        N.check(N.lib().gbcodec_decode_argmax_f32(_ptr(hm), B, K, H, W, mode, _ptr(coords), _ptr(maxvals),
                                                  _ptr(index), _stream(hm)), "decode_argmax")
    return coords, maxvals, index


@decode_argmax.register_fake
def _(hm, mode):
    B, K = hm.shape[0], hm.shape[1]
    return hm.new_empty((B, K, 2)), hm.new_empty((B, K)), hm.new_empty((B, K), dtype=torch.int32)


@torch.library.custom_op(f"{_NS}::refine_centroid", mutates_args=())
def refine_centroid(hm: Tensor, coords: Tensor, window: int) -> Tensor:
    B, K, H, W = hm.shape
    hm = _cuda_f32("heatmaps", hm)
    coords = _cuda_f32("coords", coords, (B, K, 2))
    out = torch.empty_like(coords)
    with _on_device(hm.device):
        N.check(N.lib().gbcodec_refine_centroid_f32(_ptr(hm), _ptr(coords), B, K, H, W, window, _ptr(out), _stream(hm)),
                "refine_centroid")
    return out


@refine_centroid.register_fake
def _(hm, coords, window):
    return torch.empty_like(coords)


# --------------------------------------------------------------------------- loss
_DESC_CACHE: dict = {}


def _desc(hm: Tensor, in_w, in_h, lambdas, target_sigma, encode_sigma, utw, pairs_flat):
    """struct gbcodec_loss_desc for this call.  Filling the ctypes structure (6 lambdas, up to 32 limb pairs) takes ~22 us
    of host time and a training loop asks for the same one twice per step (forward, backward): memoised by value.  The
    library only reads it."""
    B, K, H, W = hm.shape
    key = (B, K, H, W, in_w, in_h, tuple(lambdas), target_sigma, encode_sigma, bool(utw), tuple(pairs_flat))
    d = _DESC_CACHE.get(key)
    if d is None:
        pairs = [(int(pairs_flat[i]), int(pairs_flat[i + 1])) for i in range(0, len(pairs_flat), 2)]
        d = N.make_loss_desc(B, K, H, W, in_w, in_h, lambdas, target_sigma, encode_sigma, utw, pairs)
        if len(_DESC_CACHE) >= 64:
            _DESC_CACHE.clear()
        _DESC_CACHE[key] = d
    return d


def _workspace(hm: Tensor) -> Tensor:
    B, K, H, W = hm.shape
    nbytes = N.lib().gbcodec_loss_workspace_bytes(B, K, H, W)
    return torch.empty(nbytes, dtype=torch.uint8, device=hm.device)


@torch.library.custom_op(f"{_NS}::loss_denominators", mutates_args=())
def loss_denominators(weight: Tensor, gt_kps: Tensor, target_given: bool, H: int, W: int, in_w: float, in_h: float,
                      encode_sigma: float, pairs: List[int]) -> Tensor:
    B, K = gt_kps.shape[0], gt_kps.shape[1]
    weight = _cuda_f32("target_weight", weight.reshape(B, K))
    gt_kps = _cuda_f32("gt_keypoints", gt_kps, (B, K, 2))
    prs = [(int(pairs[i]), int(pairs[i + 1])) for i in range(0, len(pairs), 2)]
    desc = N.make_loss_desc(B, K, H, W, in_w, in_h, [0.0] * 6, encode_sigma, encode_sigma, True, prs)
    out = torch.empty(2, dtype=torch.float32, device=weight.device)
    nbytes = N.lib().gbcodec_loss_workspace_bytes(B, K, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    with _on_device(weight.device):
        N.check(N.lib().gbcodec_loss_denominators_f32(desc, _ptr(weight), _ptr(gt_kps), int(target_given), _ptr(out),
                                                      _ptr(ws), nbytes, _stream(weight)), "loss_denominators")
    return out


@loss_denominators.register_fake
def _(weight, gt_kps, target_given, H, W, in_w, in_h, encode_sigma, pairs):
    return weight.new_empty(2)


@torch.library.custom_op(f"{_NS}::fusion_loss", mutates_args=())
def fusion_loss(hm: Tensor, off: Tensor, var: Optional[Tensor], target: Optional[Tensor], weight: Tensor, gt_kps: Tensor,
                denoms: Optional[Tensor], grad_scale: Optional[Tensor],
                in_w: float, in_h: float, lambdas: List[float], target_sigma: float, encode_sigma: float,
                use_target_weight: bool, pairs: List[int], with_grads: bool,
                with_decode: bool, alpha_param: Optional[Tensor], fusion_weight: Optional[Tensor], radius: int,
                decode_flags: int, peer_ctx: int = 0, peer_defer: bool = False
                ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> losses7, grad_hm, grad_off, grad_var, coords, scores, global_denoms, workspace (empty tensors for what was
    not asked; the workspace holds the pass's per-tile weights / patch geometry / normalisers and is what the backward
    reuses).  peer_ctx: address of a connected gbcodec peer context (sharded.PeerExchange) — the normalisers and
    the losses are then exchanged with the other ranks inside the kernels, over NVLink peer memory.  With `denoms`
    given as well (the GLOBAL sums, prefetched with `peer_denominators`) nothing is exchanged in front of the tile
    kernel, and with `peer_defer` the step only publishes its loss terms (losses7 = this rank's share; add the shares up
    later with `peer_collect_losses`)."""
    B, K, H, W = hm.shape
    hm = _cuda_f32("heatmaps", hm)
    off = _device_readable_f32("offsets", off, (B, K, 2, H, W))
    if var is not None:
        var = _cuda_f32("variances", var, (B, K, H, W))
    if target is not None:
        target = _cuda_f32("target_heatmaps", target, (B, K, H, W))
    weight = _cuda_f32("target_weight", weight.reshape(B, K), (B, K))
    gt_kps = _cuda_f32("gt_keypoints", gt_kps, (B, K, 2))
    if denoms is not None:
        denoms = _cuda_f32("denominators", denoms.reshape(2), (2,))
    grad_scale = _scalar("grad_scale", grad_scale, hm)
    dev = hm.device
    desc = _desc(hm, in_w, in_h, lambdas, target_sigma, encode_sigma, use_target_weight, pairs)
    losses = torch.empty(7, dtype=torch.float32, device=dev)
    empty = lambda: torch.empty(0, dtype=torch.float32, device=dev)
    if with_grads:
        ghm, goff = torch.empty_like(hm), torch.empty(off.shape, dtype=torch.float32, device=dev)
        gvar = torch.empty_like(var) if var is not None else empty()
    else:
        ghm, goff, gvar = empty(), empty(), empty()
    ws = _workspace(hm)
    L = N.lib()
    common = (desc, _ptr(hm), _ptr(off), _ptr(var), _ptr(target), _ptr(weight), _ptr(gt_kps), _ptr(denoms), _ptr(grad_scale),
              _ptr(losses), _ptr(ghm) if with_grads else None, _ptr(goff) if with_grads else None,
              _ptr(gvar) if (with_grads and var is not None) else None)
    den_out = empty()
    with _on_device(dev):
        if with_decode:
            coords = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
            scores = torch.empty((B, K), dtype=torch.float32, device=dev)
            alpha_param = _scalar("alpha", alpha_param, hm)
            fusion_weight = _scalar("fusion_weight", fusion_weight, hm)
        else:
            coords, scores = empty(), empty()
        if peer_ctx:
            if peer_defer and denoms is None:
                raise RuntimeError("gbcodec: deferred losses need the prefetched global `denominators`")
            den_out = torch.empty(2, dtype=torch.float32, device=dev) if denoms is None else empty()      # an output must not alias an input
            N.check(L.gbcodec_fusion_step_sharded_f32(
                desc, _ptr(hm), _ptr(off), _ptr(var), _ptr(target), _ptr(weight), _ptr(gt_kps), _ptr(grad_scale), *common[9:],
                _ptr(alpha_param) if with_decode else None, _ptr(fusion_weight) if with_decode else None, radius, decode_flags,
                _ptr(coords) if with_decode else None, _ptr(scores) if with_decode else None,
                _ptr(den_out) if denoms is None else None, _ptr(denoms), int(peer_defer),
                _ptr(ws), ws.numel(), N._P(peer_ctx), _stream(hm)), "fusion_step_sharded")
        elif with_decode:
            N.check(L.gbcodec_fusion_step_f32(*common, _ptr(alpha_param), _ptr(fusion_weight), radius, decode_flags,
                                              _ptr(coords), _ptr(scores), _ptr(ws), ws.numel(), _stream(hm)), "fusion_step")
        else:
            N.check(L.gbcodec_fusion_loss_f32(*common, _ptr(ws), ws.numel(), _stream(hm)), "fusion_loss")
    return losses, ghm, goff, gvar, coords, scores, den_out, ws


def fusion_step_into(hm: Tensor, off: Tensor, var: Optional[Tensor], weight: Tensor, gt_kps: Tensor, denoms: Optional[Tensor],
                     in_w: float, in_h: float, lambdas: List[float], target_sigma: float, encode_sigma: float,
                     use_target_weight: bool, pairs: List[int], alpha_param: Tensor, fusion_weight: Tensor, radius: int,
                     decode_flags: int, losses_out: Tensor, grads_out: Optional[Tuple[Tensor, Tensor, Optional[Tensor]]],
                     coords_out: Tensor, scores_out: Tensor, workspace: Tensor) -> None:
    """gbcodec_fusion_step_f32 writing into buffers the caller owns (no allocation, no autograd): the chunked host-buffer
    pipeline (host_step.HostCodecStep) hands in slices of its persistent (B, ...) result and gradient tensors, so every
    chunk's gradients stay on the device for the consumer.  Targets are generated on the fly; all outputs must be
    contiguous float32 CUDA tensors of the shapes gbcodec_fusion_step_f32 documents."""
    B, K, H, W = hm.shape
    hm = _cuda_f32("heatmaps", hm)
    off = _device_readable_f32("offsets", off, (B, K, 2, H, W))
    if var is not None:
        var = _cuda_f32("variances", var, (B, K, H, W))
    for name, t, shape in (("losses_out", losses_out, (7,)), ("coords_out", coords_out, (B, K, 2)), ("scores_out", scores_out, (B, K))):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape):
            raise RuntimeError(f"gbcodec: `{name}` must be a contiguous float32 CUDA tensor of shape {shape}")
    ghm = goff = gvar = None
    if grads_out is not None:
        ghm, goff, gvar = grads_out
        for name, t, shape in (("grad_hm", ghm, (B, K, H, W)), ("grad_off", goff, (B, K, 2, H, W)), ("grad_var", gvar, (B, K, H, W))):
            if t is None and name == "grad_var" and var is None:
                continue
            if t is None or not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape):
                raise RuntimeError(f"gbcodec: `{name}` must be a contiguous float32 CUDA tensor of shape {shape}")
    desc = _desc(hm, in_w, in_h, lambdas, target_sigma, encode_sigma, use_target_weight, pairs)
    with _on_device(hm.device):
        N.check(N.lib().gbcodec_fusion_step_f32(
            desc, _ptr(hm), _ptr(off), _ptr(var), None, _ptr(weight.reshape(B, K)), _ptr(gt_kps), _ptr(denoms), None,
            _ptr(losses_out), _ptr(ghm), _ptr(goff), _ptr(gvar), _ptr(alpha_param), _ptr(fusion_weight), radius, decode_flags,
            _ptr(coords_out), _ptr(scores_out), _ptr(workspace), workspace.numel(), _stream(hm)), "fusion_step")


@fusion_loss.register_fake
def _(hm, off, var, target, weight, gt_kps, denoms, grad_scale, in_w, in_h, lambdas, target_sigma, encode_sigma,
      use_target_weight, pairs, with_grads, with_decode, alpha_param, fusion_weight, radius, decode_flags, peer_ctx=0, peer_defer=False):
    B, K = hm.shape[0], hm.shape[1]
    e = lambda: hm.new_empty(0)
    return (hm.new_empty(7),
            torch.empty_like(hm) if with_grads else e(), torch.empty_like(off) if with_grads else e(),
            torch.empty_like(var) if (with_grads and var is not None) else e(),
            hm.new_empty((B, K, 2)) if with_decode else e(), hm.new_empty((B, K)) if with_decode else e(),
            hm.new_empty(2) if (peer_ctx and denoms is None) else e(), hm.new_empty(0, dtype=torch.uint8))


def _backward_call(half: bool, g7: Tensor, ghm: Tensor, goff: Tensor, gvar: Optional[Tensor], stored: bool,
                   hm: Tensor, off: Tensor, var: Optional[Tensor], target: Optional[Tensor], weight: Tensor, gt_kps: Tensor,
                   denoms: Optional[Tensor], assumed: Optional[Tensor], scalars, held: Optional[Tensor], held_valid: bool,
                   ws: Optional[Tensor]) -> None:
    """gbcodec_fusion_loss_backward_f32 / _f16 through ctypes.  `held`: 6 device floats that travel with the stored
    gradients (which upstream factor each term's share of them carries now); `ws`: the forward's workspace."""
    B, K, H, W = hm.shape
    desc = _desc(hm, *scalars)
    from_forward = ws is not None and ws.numel() > 0
    if not from_forward:
        ws = _workspace(hm)
    with _on_device(hm.device):
        if half:
            N.check(N.lib().gbcodec_fusion_loss_backward_f16(
                desc, _ptr(hm), _ptr(off), _ptr(var), _ptr(target), _ptr(weight), _ptr(gt_kps), _ptr(denoms), _ptr(assumed),
                int(stored), _ptr(g7), _ptr(ghm), _ptr(goff), _ptr(gvar), _ptr(held), int(held_valid), int(from_forward),
                _ptr(ws), ws.numel(), _stream(hm)), "fusion_loss_backward_f16")
        else:
            N.check(N.lib().gbcodec_fusion_loss_backward_f32(
                desc, _ptr(hm), _ptr(off), _ptr(var), _ptr(target), _ptr(weight), _ptr(gt_kps), _ptr(denoms), _ptr(assumed),
                _ptr(g7), _ptr(ghm), _ptr(goff), _ptr(gvar), _ptr(held), int(held_valid), int(from_forward),
                _ptr(ws), ws.numel(), _stream(hm)), "fusion_loss_backward")


@torch.library.custom_op(f"{_NS}::fusion_loss_backward", mutates_args=("grad_hm", "grad_off", "grad_var"))
def fusion_loss_backward(grad_losses: Tensor, grad_hm: Tensor, grad_off: Tensor, grad_var: Optional[Tensor],
                         hm: Tensor, off: Tensor, var: Optional[Tensor], target: Optional[Tensor], weight: Tensor,
                         gt_kps: Tensor, denoms: Optional[Tensor], grad_scale: Optional[Tensor],
                         in_w: float, in_h: float, lambdas: List[float], target_sigma: float, encode_sigma: float,
                         use_target_weight: bool, pairs: List[int]) -> None:
    """The stored gradients (written by fusion_loss for d(total) = grad_scale) brought to the upstream vector
    `grad_losses` (7): nothing / an in-place rescale / a recompute, decided on the device.  One call per forward: the
    autograd rule of fusion_loss (which may be walked several times) keeps the `held` state of include/gbcodec.h itself."""
    B, K = hm.shape[0], hm.shape[1]
    g7 = _cuda_f32("grad_losses", grad_losses.reshape(7), (7,))
    _backward_call(False, g7, grad_hm, grad_off, grad_var, True, hm, off, var, target, weight.reshape(B, K), gt_kps, denoms,
                   grad_scale, (in_w, in_h, lambdas, target_sigma, encode_sigma, use_target_weight, pairs), None, False, None)


def _loss_setup_context(ctx, inputs, output):
    (hm, off, var, target, weight, gt_kps, denoms, grad_scale, in_w, in_h, lambdas, target_sigma, encode_sigma,
     utw, pairs, with_grads, with_decode, alpha_param, fusion_weight, radius, decode_flags, peer_ctx, peer_defer) = inputs
    losses, ghm, goff, gvar, coords, scores, den_out, ws = output
    if peer_ctx and denoms is None:
        denoms = den_out          # the backward re-uses the global normalisers the forward exchanged
    ctx.with_grads = with_grads
    ctx.has_var = var is not None
    # Plain attributes, not save_for_backward: the backward adjusts the stored gradients in place, and `held` (allocated
    # by the first backward) records which upstream factors they carry afterwards.
    ctx.stash = (ghm, goff, gvar if var is not None else None)
    ctx.ws = ws
    ctx.held = None
    B, K = hm.shape[0], hm.shape[1]
    contig = lambda t: None if t is None else t.detach().contiguous()
    ctx.tensors = (contig(hm), contig(off), contig(var), contig(target), weight.detach().reshape(B, K).contiguous(),
                   contig(gt_kps), contig(denoms), _scalar("grad_scale", grad_scale, hm))
    ctx.scalars = (in_w, in_h, list(lambdas), target_sigma, encode_sigma, utw, list(pairs))
    ctx.mark_non_differentiable(ghm, goff, gvar, coords, scores, den_out, ws)
    ctx.set_materialize_grads(False)


def _loss_backward(ctx, g_losses, *_unused):
    n_in = 23
    none = [None] * n_in
    if g_losses is None:
        return tuple(none)
    if not ctx.with_grads:
        raise RuntimeError("gbcodec::fusion_loss was run with with_grads=False; its output is not differentiable")
    hm, off, var, target, weight, gt_kps, denoms, grad_scale = ctx.tensors
    if ctx.stash is not None:
        ghm, goff, gvar = ctx.stash
        held_valid = ctx.held is not None
        if not held_valid:
            ctx.held = torch.empty(6, dtype=torch.float32, device=hm.device)
    else:
        # the stored gradients left with an earlier backward through this graph: fresh buffers, computed from the inputs
        # (a `held` state of NaN matches no upstream vector -> the device-side plan is "recompute")
        ghm, goff = torch.empty_like(hm), torch.empty_like(off)
        gvar = torch.empty_like(var) if var is not None else None
        ctx.held = torch.full((6,), float("nan"), dtype=torch.float32, device=hm.device)
        held_valid = True
    g7 = g_losses.detach().reshape(7).to(torch.float32).contiguous()
    # straight through ctypes: this runs inside the autograd engine, where the dispatcher's bookkeeping for a
    # mutating custom op (tens of microseconds of host time) buys nothing
    _backward_call(False, g7, ghm, goff, gvar, True, hm, off, var, target, weight, gt_kps, denoms, grad_scale, ctx.scalars,
                   ctx.held, held_valid, ctx.ws)
    # Ownership of the three tensors passes to autograd: with no other reference left, AccumulateGrad adopts them as the
    # leaves' .grad instead of copying 856 MB at B = 1024 (0.3 ms — more than the whole step).
    ctx.stash = None
    none[0], none[1] = ghm, goff
    none[2] = gvar
    del ghm, goff, gvar
    return tuple(none)


fusion_loss.register_autograd(_loss_backward, setup_context=_loss_setup_context)


# --------------------------------------------------------------------------- eager fast path
# torch.library's dispatch of a 23-argument custom op costs ~0.25 ms of host time per call (schema normalisation,
# fill_defaults, the dynamo-disable wrappers) — as much as the whole step takes on the GPU.  The modules therefore call
# the same implementation and the same autograd rule through a plain autograd.Function when nothing is being traced;
# the registered ops stay what torch.compile / export see.
class _FusionLossDirect(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *args):
        out = fusion_loss._init_fn(*args)
        _loss_setup_context(ctx, args, out)
        return out

    @staticmethod
    def backward(ctx, *grads):
        return _loss_backward(ctx, *grads)


class _FusionLossF16Direct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *args):
        out = fusion_loss_f16._init_fn(*args)
        _loss_f16_setup(ctx, args, out)
        return out

    @staticmethod
    def backward(ctx, *grads):
        return _loss_f16_backward(ctx, *grads)


def fusion_loss_eager(*args):
    """`fusion_loss` with every argument given positionally (23 of them), without the dispatcher."""
    if torch.compiler.is_compiling():
        return fusion_loss(*args)
    return _FusionLossDirect.apply(*args)


def fusion_loss_f16_eager(*args):
    if torch.compiler.is_compiling():
        return fusion_loss_f16(*args)
    return _FusionLossF16Direct.apply(*args)


def peer_denominators(weight: Tensor, gt_kps: Tensor, target_given: bool, H: int, W: int, in_w: float, in_h: float,
                      encode_sigma: float, pairs: List[int], peer_ctx: int, out: Optional[Tensor] = None) -> Tensor:
    """gbcodec_peer_denominators_f32: this rank's normaliser sums into every peer's mailbox, the GLOBAL sums out (2 floats
    on the device).  Issue it for the NEXT batch on a side stream while the current step runs (the sums depend on the
    visibility flags and keypoints only); every rank must call it once per step, in step order."""
    B, K = gt_kps.shape[0], gt_kps.shape[1]
    weight = _cuda_f32("target_weight", weight.reshape(B, K))
    gt_kps = _cuda_f32("gt_keypoints", gt_kps, (B, K, 2))
    prs = [(int(pairs[i]), int(pairs[i + 1])) for i in range(0, len(pairs), 2)]
    desc = N.make_loss_desc(B, K, H, W, in_w, in_h, [0.0] * 6, encode_sigma, encode_sigma, True, prs)
    out = torch.empty(2, dtype=torch.float32, device=weight.device) if out is None else out
    nbytes = N.lib().gbcodec_loss_workspace_bytes(B, K, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    with _on_device(weight.device):
        N.check(N.lib().gbcodec_peer_denominators_f32(desc, _ptr(weight), _ptr(gt_kps), int(target_given), _ptr(out), _ptr(ws), nbytes,
                                                      N._P(peer_ctx), _stream(weight)), "peer_denominators")
    return out


def peer_collect_losses(peer_ctx: int, device, steps_back: int = 0, out: Optional[Tensor] = None) -> Tensor:
    """gbcodec_peer_collect_losses_f32: the ranks' loss terms of the sharded step made `steps_back` calls ago, added in
    rank order -> the 7 global losses (on the device; read them when you log)."""
    out = torch.empty(7, dtype=torch.float32, device=device) if out is None else out
    with _on_device(out.device):
        N.check(N.lib().gbcodec_peer_collect_losses_f32(N._P(peer_ctx), int(steps_back), _ptr(out), _stream(out)), "peer_collect_losses")
    return out


# --------------------------------------------------------------------------- variance branch tail: Softplus -> mean_N
@torch.library.custom_op(f"{_NS}::softplus_mean", mutates_args=())
def softplus_mean(raw: Tensor) -> Tensor:
    """mean over each (H,W) tile of softplus(raw): (B,K,H,W) -> (B,K).  The variance branch of the fusion head ends in
    Softplus (fusion_head.py:245-251) and the loss reads only this mean (:467-478)."""
    B, K, H, W = raw.shape
    raw = _cuda_f32("raw_variances", raw)
    out = torch.empty((B, K), dtype=torch.float32, device=raw.device)
    with _on_device(raw.device):
        N.check(N.lib().gbcodec_softplus_mean_f32(_ptr(raw), _ptr(out), B, K, H, W, _stream(raw)), "softplus_mean")
    return out


@softplus_mean.register_fake
def _(raw):
    return raw.new_empty(raw.shape[:2])


@torch.library.custom_op(f"{_NS}::softplus_mean_backward", mutates_args=())
def softplus_mean_backward(raw: Tensor, grad_mean: Tensor) -> Tensor:
    B, K, H, W = raw.shape
    raw = _cuda_f32("raw_variances", raw)
    grad_mean = _cuda_f32("grad_mean", grad_mean.reshape(B, K), (B, K))
    out = torch.empty_like(raw)
    with _on_device(raw.device):
        N.check(N.lib().gbcodec_softplus_mean_backward_f32(_ptr(raw), _ptr(grad_mean), _ptr(out), B, K, H, W, _stream(raw)),
                "softplus_mean_backward")
    return out


@softplus_mean_backward.register_fake
def _(raw, grad_mean):
    return torch.empty_like(raw)


softplus_mean.register_autograd(lambda ctx, g: softplus_mean_backward(ctx.saved_tensors[0], g),
                                setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))


def pairs_flat(pairs: Sequence[Tuple[int, int]]) -> List[int]:
    out: List[int] = []
    for a, b in pairs:
        out += [int(a), int(b)]
    return out


# =========================================================================== Gen-B family
@torch.library.custom_op(f"{_NS}::encode_mode", mutates_args=())
def encode_mode(kps: Tensor, vis: Tensor, H: int, W: int, in_w: float, in_h: float, sigma: float, mode: int
                ) -> Tuple[Tensor, Tensor]:
    """mode: N.ENCODE_PATCH | N.ENCODE_PATCH_CLIPPED | N.ENCODE_DENSE -> target (B,K,H,W), weight (B,K,1)."""
    B, K = kps.shape[0], kps.shape[1]
    kps = _cuda_f32("keypoints", kps, (B, K, 2))
    vis = _cuda_f32("visible", vis.reshape(B, K), (B, K))
    target = torch.empty((B, K, H, W), dtype=torch.float32, device=kps.device)
    weight = torch.empty((B, K, 1), dtype=torch.float32, device=kps.device)
    with _on_device(kps.device):
        N.check(N.lib().gbcodec_encode_mode_f32(_ptr(kps), _ptr(vis), _ptr(target), _ptr(weight), B, K, H, W,
                                                in_w, in_h, sigma, mode, _stream(kps)), "encode_mode")
    return target, weight


@encode_mode.register_fake
def _(kps, vis, H, W, in_w, in_h, sigma, mode):
    B, K = kps.shape[0], kps.shape[1]
    return kps.new_empty((B, K, H, W)), kps.new_empty((B, K, 1))


@torch.library.custom_op(f"{_NS}::postprocess", mutates_args=())
def postprocess(hm: Tensor, regression: Optional[Tensor], center: Optional[Tensor], scale: Optional[Tensor],
                argmax_mode: int, scale_to_image: bool, image_size: float, refine_window: int,
                filter_low: bool, threshold: float, transform: bool, input_w: float, input_h: float
                ) -> Tuple[Tensor, Tensor, Tensor]:
    """utils/postprocess.py:296-340 in one kernel -> preds (B,K,2), maxvals (B,K), mask (B,K)."""
    B, K, H, W = hm.shape
    hm = _cuda_f32("heatmaps", hm)
    dev = hm.device
    if regression is not None:
        regression = _cuda_f32("regression_coords", regression, (B, K, 2))
    if center is not None:
        center = _cuda_f32("center", center, (B, 2))
    if scale is not None:
        scale = _cuda_f32("scale", scale, (B, 2))
    d = N.PostprocessDesc(B, K, H, W, argmax_mode, int(scale_to_image), image_size, refine_window, int(filter_low),
                          threshold, int(transform), input_w, input_h)
    preds = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    maxvals = torch.empty((B, K), dtype=torch.float32, device=dev)
    mask = torch.empty((B, K), dtype=torch.float32, device=dev)
    ws = torch.empty(16, dtype=torch.uint8, device=dev)
    with _on_device(dev):
        N.check(N.lib().gbcodec_postprocess_f32(d, _ptr(hm), _ptr(regression), _ptr(center), _ptr(scale), _ptr(preds),
                                                _ptr(maxvals), _ptr(mask), _ptr(ws), _stream(hm)), "postprocess")
    return preds, maxvals, mask


@postprocess.register_fake
def _(hm, regression, center, scale, argmax_mode, scale_to_image, image_size, refine_window, filter_low, threshold,
      transform, input_w, input_h):
    B, K = hm.shape[0], hm.shape[1]
    return hm.new_empty((B, K, 2)), hm.new_empty((B, K)), hm.new_empty((B, K))


@torch.library.custom_op(f"{_NS}::coords_to_image", mutates_args=())
def coords_to_image(coords: Tensor, center: Tensor, scale: Tensor, H: int, W: int, in_w: float, in_h: float) -> Tensor:
    """validate.py:102-119 — heatmap px -> input px -> original image."""
    B, K = coords.shape[0], coords.shape[1]
    coords = _cuda_f32("coords", coords, (B, K, 2))
    center = _cuda_f32("center", center, (B, 2))
    scale = _cuda_f32("scale", scale, (B, 2))
    out = torch.empty_like(coords)
    with _on_device(coords.device):
        N.check(N.lib().gbcodec_coords_to_image_f32(_ptr(coords), _ptr(center), _ptr(scale), B, K, H, W, in_w, in_h,
                                                    _ptr(out), _stream(coords)), "coords_to_image")
    return out


@coords_to_image.register_fake
def _(coords, center, scale, H, W, in_w, in_h):
    return torch.empty_like(coords)


def _combined_desc(B, K, H, W, norm_batch, terms, heat_crit, coord_crit, utw, heat_scale, lam_var, lam_mean, weights):
    return N.CombinedDesc(B, K, H, W, norm_batch, terms, heat_crit, coord_crit, int(utw), heat_scale, lam_var, lam_mean,
                          weights[0], weights[1], weights[2])


def _combined_shapes(pred, coords, refined, target_coords):
    if pred is not None:
        return tuple(pred.shape)
    c = coords if coords is not None else refined
    return (c.shape[0], c.shape[1], 0, 0)


@torch.library.custom_op(f"{_NS}::combined_loss", mutates_args=())
def combined_loss(pred: Optional[Tensor], target: Optional[Tensor], weight: Optional[Tensor],
                  coords: Optional[Tensor], refined: Optional[Tensor], target_coords: Optional[Tensor],
                  grad_scale: Optional[Tensor], norm_batch: int, heat_crit: int, coord_crit: int,
                  use_target_weight: bool, heat_scale: float, lam_var: float, lam_mean: float, weights: List[float],
                  morph: bool, with_grads: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """models/losses.py:205-290 -> losses5 (heatmap, morph, regression, refined, total), grad_pred, grad_coords,
    grad_refined (empty tensors for what was not asked / not present).  `pred` may be float16 (the head's heatmaps
    under autocast; every other tensor float32): gbcodec_combined_loss_f16 up-casts it where it is read and leaves
    grad_pred in float16, rounded once after the upstream factor `grad_scale`."""
    B, K, H, W = _combined_shapes(pred, coords, refined, target_coords)
    some = pred if pred is not None else (coords if coords is not None else refined)
    dev = some.device
    terms = 0
    half = pred is not None and pred.dtype == torch.float16
    if pred is not None:
        pred = _cuda_f16_or_f32("pred_heatmaps", pred, (B, K, H, W))
        target = _cuda_f32("target_heatmaps", target, (B, K, H, W))
        terms |= N.TERM_HEATMAP | (N.TERM_MORPH if morph else 0)
    if coords is not None:
        coords = _cuda_f32("pred_coords", coords, (B, K, 2))
        terms |= N.TERM_REGRESSION
    if refined is not None:
        refined = _cuda_f32("refined_coords", refined, (B, K, 2))
        terms |= N.TERM_REFINED
    if target_coords is not None:
        target_coords = _cuda_f32("target_coords", target_coords, (B, K, 2))
    if weight is not None:
        weight = _cuda_f32("target_weight", weight.reshape(B, K), (B, K))
    grad_scale = _scalar("grad_scale", grad_scale, some)
    desc = _combined_desc(B, K, H, W, norm_batch, terms, heat_crit, coord_crit, use_target_weight, heat_scale, lam_var,
                          lam_mean, weights)
    losses = torch.empty(5, dtype=torch.float32, device=dev)
    empty = lambda: torch.empty(0, dtype=torch.float32, device=dev)
    gp = torch.empty_like(pred) if (with_grads and pred is not None) else torch.empty(0, dtype=torch.float16 if half else torch.float32, device=dev)
    gc = torch.empty_like(coords) if (with_grads and coords is not None) else empty()
    gr = torch.empty_like(refined) if (with_grads and refined is not None) else empty()
    nbytes = N.lib().gbcodec_combined_workspace_bytes(B, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    opt = lambda t: _ptr(t) if t.numel() else None
    entry = N.lib().gbcodec_combined_loss_f16 if half else N.lib().gbcodec_combined_loss_f32
    with _on_device(dev):
        N.check(entry(desc, _ptr(pred), _ptr(target), _ptr(weight), _ptr(coords), _ptr(refined), _ptr(target_coords),
                      _ptr(grad_scale), _ptr(losses), opt(gp), opt(gc), opt(gr), _ptr(ws), nbytes, _stream(some)), "combined_loss")
    return losses, gp, gc, gr


@combined_loss.register_fake
def _(pred, target, weight, coords, refined, target_coords, grad_scale, norm_batch, heat_crit, coord_crit,
      use_target_weight, heat_scale, lam_var, lam_mean, weights, morph, with_grads):
    some = pred if pred is not None else (coords if coords is not None else refined)
    e = lambda: some.new_empty(0, dtype=torch.float32)
    g = lambda t: torch.empty_like(t) if (with_grads and t is not None) else e()
    gp = torch.empty_like(pred) if (with_grads and pred is not None) else some.new_empty(0)
    return some.new_empty(5, dtype=torch.float32), gp, g(coords), g(refined)


@torch.library.custom_op(f"{_NS}::combined_loss_backward", mutates_args=("grad_pred", "grad_coords", "grad_refined"))
def combined_loss_backward(grad_losses: Tensor, grad_pred: Tensor, grad_coords: Tensor, grad_refined: Tensor,
                           pred: Optional[Tensor], target: Optional[Tensor], weight: Optional[Tensor],
                           coords: Optional[Tensor], refined: Optional[Tensor], target_coords: Optional[Tensor],
                           grad_scale: Optional[Tensor], norm_batch: int, heat_crit: int, coord_crit: int,
                           use_target_weight: bool, heat_scale: float, lam_var: float, lam_mean: float,
                           weights: List[float], morph: bool) -> None:
    B, K, H, W = _combined_shapes(pred, coords, refined, target_coords)
    some = pred if pred is not None else (coords if coords is not None else refined)
    terms = 0
    if pred is not None:
        terms |= N.TERM_HEATMAP | (N.TERM_MORPH if morph else 0)
    if coords is not None:
        terms |= N.TERM_REGRESSION
    if refined is not None:
        terms |= N.TERM_REFINED
    if weight is not None:
        weight = weight.reshape(B, K)
    desc = _combined_desc(B, K, H, W, norm_batch, terms, heat_crit, coord_crit, use_target_weight, heat_scale, lam_var,
                          lam_mean, weights)
    g5 = _cuda_f32("grad_losses", grad_losses.reshape(5), (5,))
    nbytes = N.lib().gbcodec_combined_workspace_bytes(B, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=some.device)
    opt = lambda t: _ptr(t) if t.numel() else None
    half = pred is not None and pred.dtype == torch.float16
    if half and grad_pred.dtype != torch.float16:
        raise RuntimeError("gbcodec: float16 predictions take a float16 grad_pred")
    entry = N.lib().gbcodec_combined_loss_backward_f16 if half else N.lib().gbcodec_combined_loss_backward_f32
    with _on_device(some.device):
        N.check(entry(
            desc, _ptr(pred), _ptr(target), _ptr(weight), _ptr(coords), _ptr(refined), _ptr(target_coords), _ptr(grad_scale),
            _ptr(g5), opt(grad_pred), opt(grad_coords), opt(grad_refined), _ptr(ws), nbytes, _stream(some)),
            "combined_loss_backward")


def _combined_setup_context(ctx, inputs, output):
    (pred, target, weight, coords, refined, target_coords, grad_scale, norm_batch, heat_crit, coord_crit, utw, heat_scale,
     lam_var, lam_mean, weights, morph, with_grads) = inputs
    losses, gp, gc, gr = output
    ctx.with_grads = with_grads
    ctx.stash = (gp, gc, gr)
    ctx.tensors = tuple(None if t is None else t.detach() for t in (pred, target, weight, coords, refined, target_coords, grad_scale))
    ctx.scalars = (norm_batch, heat_crit, coord_crit, utw, heat_scale, lam_var, lam_mean, list(weights), morph)
    ctx.set_materialize_grads(False)


def _combined_backward(ctx, g_losses, g_gp, g_gc, g_gr):
    none = [None] * 17
    if g_losses is None:
        return tuple(none)
    if not ctx.with_grads:
        raise RuntimeError("gbcodec::combined_loss was run with with_grads=False; its output is not differentiable")
    pred, target, weight, coords, refined, target_coords, grad_scale = ctx.tensors
    contig = lambda t: None if t is None else t.contiguous()
    if ctx.stash is not None:
        gp, gc, gr = ctx.stash
        assumed = grad_scale
    else:
        # The stored gradients left with an earlier backward through this graph (and that backward may have recomputed
        # them for another upstream vector): fresh buffers, and an "assumed" scale of NaN, which no upstream matches, so
        # that the device-side plan is "recompute" — never stale gradients.
        some = pred if pred is not None else (coords if coords is not None else refined)
        e = lambda: torch.empty(0, dtype=torch.float32, device=some.device)
        gp = torch.empty_like(pred) if pred is not None else e()
        gc = torch.empty_like(coords) if coords is not None else e()
        gr = torch.empty_like(refined) if refined is not None else e()
        assumed = torch.full((1,), float("nan"), dtype=torch.float32, device=some.device)
    combined_loss_backward._init_fn(g_losses.detach().contiguous(), gp, gc, gr, contig(pred), contig(target), contig(weight), contig(coords),
                                    contig(refined), contig(target_coords), assumed, *ctx.scalars)
    ctx.stash = None                    # ownership passes to autograd (AccumulateGrad adopts the tensors instead of copying them)
    none[0] = gp if pred is not None else None
    none[3] = gc if coords is not None else None
    none[4] = gr if refined is not None else None
    del gp, gc, gr
    return tuple(none)


combined_loss.register_autograd(_combined_backward, setup_context=_combined_setup_context)


# --------------------------------------------------------------------------- plain heatmap head, one pass
@torch.library.custom_op(f"{_NS}::heatmap_step", mutates_args=())
def heatmap_step(hm: Tensor, target: Optional[Tensor], weight: Optional[Tensor], gt_kps: Optional[Tensor],
                 in_w: float, in_h: float, sigma: float, use_target_weight: bool, norm_batch: int,
                 grad_scale: Optional[Tensor], with_grads: bool, with_decode: bool, argmax_mode: int
                 ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """KeypointMSELoss fwd (+ bwd) [+ decode_heatmaps] in one pass -> loss (0-dim), grad_hm, coords (B,K,2), maxvals (B,K);
    target None: tiles and weights are generated in the kernel from gt_kps and weight (= visibility)."""
    B, K, H, W = hm.shape
    hm = _cuda_f32("heatmaps", hm)
    dev = hm.device
    if target is not None:
        target = _cuda_f32("target_heatmaps", target, (B, K, H, W))
    if weight is not None:
        weight = _cuda_f32("target_weight", weight.reshape(B, K), (B, K))
    if gt_kps is not None:
        gt_kps = _cuda_f32("gt_keypoints", gt_kps, (B, K, 2))
    grad_scale = _scalar("grad_scale", grad_scale, hm)
    empty = lambda: torch.empty(0, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    ghm = torch.empty_like(hm) if with_grads else empty()
    coords = torch.empty((B, K, 2), dtype=torch.float32, device=dev) if with_decode else empty()
    maxvals = torch.empty((B, K), dtype=torch.float32, device=dev) if with_decode else empty()
    nbytes = N.lib().gbcodec_combined_workspace_bytes(B, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with _on_device(dev):
        N.check(N.lib().gbcodec_heatmap_step_f32(
            _ptr(hm), _ptr(target), _ptr(weight), _ptr(gt_kps), B, K, H, W, in_w, in_h, sigma, int(use_target_weight), norm_batch,
            _ptr(grad_scale), _ptr(loss), _ptr(ghm) if with_grads else None, argmax_mode,
            _ptr(coords) if with_decode else None, _ptr(maxvals) if with_decode else None, None,
            _ptr(ws), nbytes, _stream(hm)), "heatmap_step")
    return loss, ghm, coords, maxvals


@heatmap_step.register_fake
def _(hm, target, weight, gt_kps, in_w, in_h, sigma, use_target_weight, norm_batch, grad_scale, with_grads, with_decode, argmax_mode):
    B, K = hm.shape[0], hm.shape[1]
    e = lambda: hm.new_empty(0)
    return (hm.new_empty(()), torch.empty_like(hm) if with_grads else e(),
            hm.new_empty((B, K, 2)) if with_decode else e(), hm.new_empty((B, K)) if with_decode else e())


def _heatmap_step_setup(ctx, inputs, output):
    ctx.with_grads = inputs[10]
    ctx.grad_scale = inputs[9]
    ctx.ghm = output[1]
    ctx.mark_non_differentiable(output[1], output[2], output[3])      # the stored gradient, keypoints, max values
    ctx.set_materialize_grads(False)


def _heatmap_step_backward(ctx, g_loss, g_ghm, g_coords, g_maxvals):
    none = [None] * 13
    if g_loss is None:
        return tuple(none)
    if not ctx.with_grads:
        raise RuntimeError("gbcodec::heatmap_step was run with with_grads=False; its output is not differentiable")
    # the stored gradient assumes d(loss) = grad_scale (1 if absent): one scalar multiply otherwise, on the device
    ratio = g_loss if ctx.grad_scale is None else g_loss / ctx.grad_scale.reshape(())
    none[0] = ctx.ghm * ratio
    return tuple(none)


heatmap_step.register_autograd(_heatmap_step_backward, setup_context=_heatmap_step_setup)


# --------------------------------------------------------------------------- float16 maps (autocast)
HALF_TILE_SHAPES = ((64, 48), (64, 64), (96, 72), (128, 128))


def _cuda_f16_or_f32(name: str, t: Tensor, shape: Sequence[int]) -> Tensor:
    return _cuda_f16(name, t, shape) if t.dtype == torch.float16 else _cuda_f32(name, t, shape)


def _cuda_f16(name: str, t: Tensor, shape: Sequence[int]) -> Tensor:
    if not t.is_cuda or t.dtype != torch.float16 or tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"gbcodec: `{name}` must be a CUDA float16 tensor of shape {tuple(shape)}")
    return t.contiguous()


@torch.library.custom_op(f"{_NS}::fusion_loss_f16", mutates_args=())
def fusion_loss_f16(hm: Tensor, off: Tensor, var: Optional[Tensor], target: Optional[Tensor], weight: Tensor, gt_kps: Tensor,
                    denoms: Optional[Tensor], expected_upstream: Optional[Tensor], in_w: float, in_h: float, lambdas: List[float],
                    target_sigma: float, encode_sigma: float, use_target_weight: bool, pairs: List[int], with_grads: bool,
                    with_decode: bool, alpha_param: Optional[Tensor], fusion_weight: Optional[Tensor], radius: int,
                    decode_flags: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """float16 heatmaps / offsets / variances -> losses7 (float32), coords, scores, grad_hm, grad_off, grad_var (float16),
    workspace.
    With `with_grads` the pass also stores the gradients, pre-multiplied by `expected_upstream` (a 1-element float32
    device tensor: the upstream gradient of total_loss the caller expects, i.e. the loss scale; None = 1) — the
    backward keeps them if the expectation held and computes them again otherwise."""
    B, K, H, W = hm.shape
    hm = _cuda_f16("heatmaps", hm, (B, K, H, W))
    off = _cuda_f16("offsets", off, (B, K, 2, H, W))
    if var is not None:
        var = _cuda_f16("variances", var, (B, K, H, W))
    if target is not None:
        target = _cuda_f32("target_heatmaps", target, (B, K, H, W))
    weight = _cuda_f32("target_weight", weight.reshape(B, K), (B, K))
    gt_kps = _cuda_f32("gt_keypoints", gt_kps, (B, K, 2))
    if denoms is not None:
        denoms = _cuda_f32("denominators", denoms.reshape(2), (2,))
    expected_upstream = _scalar("expected_upstream", expected_upstream, hm)
    dev = hm.device
    desc = _desc(hm, in_w, in_h, lambdas, target_sigma, encode_sigma, use_target_weight, pairs)
    losses = torch.empty(7, dtype=torch.float32, device=dev)
    empty = lambda dt=torch.float32: torch.empty(0, dtype=dt, device=dev)
    coords = torch.empty((B, K, 2), dtype=torch.float32, device=dev) if with_decode else empty()
    scores = torch.empty((B, K), dtype=torch.float32, device=dev) if with_decode else empty()
    ghm = torch.empty_like(hm) if with_grads else empty(torch.float16)
    goff = torch.empty_like(off) if with_grads else empty(torch.float16)
    gvar = torch.empty_like(var) if (with_grads and var is not None) else empty(torch.float16)
    if with_decode:
        alpha_param = _scalar("alpha", alpha_param, hm)
        fusion_weight = _scalar("fusion_weight", fusion_weight, hm)
    ws = _workspace(hm)
    with _on_device(dev):
        N.check(N.lib().gbcodec_fusion_step_f16(
            desc, _ptr(hm), _ptr(off), _ptr(var), _ptr(target), _ptr(weight), _ptr(gt_kps), _ptr(denoms), _ptr(expected_upstream),
            _ptr(losses), _ptr(ghm) if with_grads else None, _ptr(goff) if with_grads else None,
            _ptr(gvar) if (with_grads and var is not None) else None,
            _ptr(alpha_param) if with_decode else None, _ptr(fusion_weight) if with_decode else None, radius, decode_flags,
            _ptr(coords) if with_decode else None, _ptr(scores) if with_decode else None, _ptr(ws), ws.numel(), _stream(hm)),
            "fusion_step_f16")
    return losses, coords, scores, ghm, goff, gvar, ws


@fusion_loss_f16.register_fake
def _(hm, off, var, target, weight, gt_kps, denoms, expected_upstream, in_w, in_h, lambdas, target_sigma, encode_sigma,
      use_target_weight, pairs, with_grads, with_decode, alpha_param, fusion_weight, radius, decode_flags):
    B, K = hm.shape[0], hm.shape[1]
    e = lambda dt=torch.float32: hm.new_empty(0, dtype=dt)
    return (hm.new_empty(7, dtype=torch.float32), hm.new_empty((B, K, 2), dtype=torch.float32) if with_decode else e(),
            hm.new_empty((B, K), dtype=torch.float32) if with_decode else e(),
            torch.empty_like(hm) if with_grads else e(torch.float16), torch.empty_like(off) if with_grads else e(torch.float16),
            torch.empty_like(var) if (with_grads and var is not None) else e(torch.float16), hm.new_empty(0, dtype=torch.uint8))


@torch.library.custom_op(f"{_NS}::fusion_loss_backward_f16", mutates_args=("grad_hm", "grad_off", "grad_var"))
def fusion_loss_backward_f16(grad_losses: Tensor, grad_hm: Tensor, grad_off: Tensor, grad_var: Optional[Tensor], stored: bool,
                             hm: Tensor, off: Tensor, var: Optional[Tensor], target: Optional[Tensor],
                             weight: Tensor, gt_kps: Tensor, denoms: Optional[Tensor], expected_upstream: Optional[Tensor],
                             in_w: float, in_h: float, lambdas: List[float], target_sigma: float, encode_sigma: float,
                             use_target_weight: bool, pairs: List[int]) -> None:
    B, K = hm.shape[0], hm.shape[1]
    g7 = _cuda_f32("grad_losses", grad_losses.reshape(7), (7,))
    _backward_call(True, g7, grad_hm, grad_off, grad_var, stored, hm, off, var, target, weight.reshape(B, K), gt_kps, denoms,
                   expected_upstream, (in_w, in_h, lambdas, target_sigma, encode_sigma, use_target_weight, pairs), None, False, None)


def _loss_f16_setup(ctx, inputs, output):
    (hm, off, var, target, weight, gt_kps, denoms, expected_upstream, in_w, in_h, lambdas, target_sigma, encode_sigma, utw, pairs,
     with_grads, with_decode, alpha_param, fusion_weight, radius, decode_flags) = inputs
    losses, coords, scores, ghm, goff, gvar, ws = output
    ctx.stored = with_grads
    ctx.stash = (ghm, goff, gvar if var is not None else None)
    ctx.ws = ws
    ctx.held = None
    # the expectation the stored gradients were built on: a private copy, the caller's tensor moves on
    ctx.expected = None if expected_upstream is None else expected_upstream.detach().to(torch.float32).reshape(1).clone()
    B, K = hm.shape[0], hm.shape[1]
    contig = lambda t: None if t is None else t.detach().contiguous()
    ctx.tensors = (contig(hm), contig(off), contig(var), contig(target), weight.detach().reshape(B, K).contiguous(), contig(gt_kps),
                   contig(denoms))
    ctx.scalars = (in_w, in_h, list(lambdas), target_sigma, encode_sigma, utw, list(pairs))
    ctx.mark_non_differentiable(coords, scores, ghm, goff, gvar, ws)
    ctx.set_materialize_grads(False)


def _loss_f16_backward(ctx, g_losses, *_unused):
    none = [None] * 21
    if g_losses is None:
        return tuple(none)
    hm, off, var, target, weight, gt_kps, denoms = ctx.tensors
    stored = ctx.stored and ctx.stash is not None
    if stored:
        ghm, goff, gvar = ctx.stash
    else:
        # nothing stored (forward under no expectation of a backward), or the stored gradients left with an earlier
        # backward through this graph: fresh buffers, always computed
        ghm, goff = torch.empty_like(hm), torch.empty_like(off)
        gvar = torch.empty_like(var) if var is not None else None
    held_valid = stored and ctx.held is not None
    if stored and not held_valid:
        ctx.held = torch.empty(6, dtype=torch.float32, device=hm.device)
    g7 = g_losses.detach().reshape(7).to(torch.float32).contiguous()
    _backward_call(True, g7, ghm, goff, gvar, stored, hm, off, var, target, weight, gt_kps, denoms, ctx.expected, ctx.scalars,
                   ctx.held if stored else None, held_valid, ctx.ws)
    ctx.stash = None                    # ownership passes to autograd (no copy in AccumulateGrad)
    none[0], none[1] = ghm, goff
    none[2] = gvar if var is not None else None
    del ghm, goff, gvar
    return tuple(none)


fusion_loss_f16.register_autograd(_loss_f16_backward, setup_context=_loss_f16_setup)


# --------------------------------------------------------------------------- variance branch given as per-tile means
@torch.library.custom_op(f"{_NS}::fusion_step_vmean", mutates_args=())
def fusion_step_vmean(hm: Tensor, off: Tensor, var_mean: Tensor, target: Optional[Tensor], weight: Tensor, gt_kps: Tensor,
                      denoms: Optional[Tensor], grad_scale: Optional[Tensor], in_w: float, in_h: float, lambdas: List[float],
                      target_sigma: float, encode_sigma: float, use_target_weight: bool, pairs: List[int], with_grads: bool,
                      with_decode: bool, alpha_param: Optional[Tensor], fusion_weight: Optional[Tensor], radius: int,
                      decode_flags: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """The fused step for a head that hands over mean_N(V) per tile instead of the variance map (16N instead of 24N bytes
    per tile) -> losses7, grad_hm, grad_off, grad_var_mean (B,K) = d(total)/d(mean_N(V)) times grad_scale, coords, scores,
    workspace.  Differentiable w.r.t. hm, off and var_mean (autograd rule below; `softplus_mean` is the other half)."""
    B, K, H, W = hm.shape
    hm = _cuda_f32("heatmaps", hm)
    off = _cuda_f32("offsets", off, (B, K, 2, H, W))
    var_mean = _cuda_f32("variance_means", var_mean.reshape(B, K), (B, K))
    if target is not None:
        target = _cuda_f32("target_heatmaps", target, (B, K, H, W))
    weight = _cuda_f32("target_weight", weight.reshape(B, K), (B, K))
    gt_kps = _cuda_f32("gt_keypoints", gt_kps, (B, K, 2))
    if denoms is not None:
        denoms = _cuda_f32("denominators", denoms.reshape(2), (2,))
    grad_scale = _scalar("grad_scale", grad_scale, hm)
    dev = hm.device
    desc = _desc(hm, in_w, in_h, lambdas, target_sigma, encode_sigma, use_target_weight, pairs)
    losses = torch.empty(7, dtype=torch.float32, device=dev)
    empty = lambda: torch.empty(0, dtype=torch.float32, device=dev)
    ghm = torch.empty_like(hm) if with_grads else empty()
    goff = torch.empty_like(off) if with_grads else empty()
    gvm = torch.empty((B, K), dtype=torch.float32, device=dev) if with_grads else empty()
    coords = torch.empty((B, K, 2), dtype=torch.float32, device=dev) if with_decode else empty()
    scores = torch.empty((B, K), dtype=torch.float32, device=dev) if with_decode else empty()
    if with_decode:
        alpha_param = _scalar("alpha", alpha_param, hm)
        fusion_weight = _scalar("fusion_weight", fusion_weight, hm)
    ws = _workspace(hm)
    opt = lambda t, on: _ptr(t) if on else None
    with _on_device(dev):
        N.check(N.lib().gbcodec_fusion_step_vmean_f32(
            desc, _ptr(hm), _ptr(off), _ptr(var_mean), _ptr(target), _ptr(weight), _ptr(gt_kps), _ptr(denoms), _ptr(grad_scale),
            _ptr(losses), opt(ghm, with_grads), opt(goff, with_grads), opt(gvm, with_grads),
            opt(alpha_param, with_decode), opt(fusion_weight, with_decode), radius, decode_flags,
            opt(coords, with_decode), opt(scores, with_decode), _ptr(ws), ws.numel(), _stream(hm)), "fusion_step_vmean")
    return losses, ghm, goff, gvm, coords, scores, ws


@fusion_step_vmean.register_fake
def _(hm, off, var_mean, target, weight, gt_kps, denoms, grad_scale, in_w, in_h, lambdas, target_sigma, encode_sigma,
      use_target_weight, pairs, with_grads, with_decode, alpha_param, fusion_weight, radius, decode_flags):
    B, K = hm.shape[0], hm.shape[1]
    e = lambda: hm.new_empty(0)
    return (hm.new_empty(7), torch.empty_like(hm) if with_grads else e(), torch.empty_like(off) if with_grads else e(),
            hm.new_empty((B, K)) if with_grads else e(), hm.new_empty((B, K, 2)) if with_decode else e(),
            hm.new_empty((B, K)) if with_decode else e(), hm.new_empty(0, dtype=torch.uint8))


def _vmean_setup(ctx, inputs, output):
    (hm, off, var_mean, target, weight, gt_kps, denoms, grad_scale, in_w, in_h, lambdas, target_sigma, encode_sigma, utw, pairs,
     with_grads, with_decode, alpha_param, fusion_weight, radius, decode_flags) = inputs
    losses, ghm, goff, gvm, coords, scores, ws = output
    ctx.with_grads = with_grads
    ctx.stash, ctx.gvm0, ctx.ws, ctx.held = (ghm, goff), gvm, ws, None
    B, K = hm.shape[0], hm.shape[1]
    contig = lambda t: None if t is None else t.detach().contiguous()
    ctx.tensors = (contig(hm), contig(off), contig(target), weight.detach().reshape(B, K).contiguous(), contig(gt_kps), contig(denoms),
                   _scalar("grad_scale", grad_scale, hm))
    ctx.scalars = (in_w, in_h, list(lambdas), target_sigma, encode_sigma, utw, list(pairs))
    ctx.mark_non_differentiable(ghm, goff, gvm, coords, scores, ws)
    ctx.set_materialize_grads(False)


def _vmean_backward(ctx, g_losses, *_unused):
    none = [None] * 21
    if g_losses is None:
        return tuple(none)
    if not ctx.with_grads:
        raise RuntimeError("gbcodec::fusion_step_vmean was run with with_grads=False; its output is not differentiable")
    hm, off, target, weight, gt_kps, denoms, grad_scale = ctx.tensors
    if ctx.stash is not None:
        ghm, goff = ctx.stash
        held_valid = ctx.held is not None
        if not held_valid:
            ctx.held = torch.empty(6, dtype=torch.float32, device=hm.device)
    else:                               # handed over by an earlier backward: recompute into fresh buffers
        ghm, goff = torch.empty_like(hm), torch.empty_like(off)
        ctx.held = torch.full((6,), float("nan"), dtype=torch.float32, device=hm.device)
        held_valid = True
    g7 = g_losses.detach().reshape(7).to(torch.float32).contiguous()
    # the heatmap / offset gradients do not see the variance branch: the map-less backward brings them to the upstream
    # vector (nothing / rescale / recompute, decided on the device); the (B,K) gradient of the means is linear in the
    # variance term's upstream factor
    _backward_call(False, g7, ghm, goff, None, True, hm, off, None, target, weight, gt_kps, denoms, grad_scale, ctx.scalars,
                   ctx.held, held_valid, ctx.ws)
    assumed = grad_scale.reshape(()) if grad_scale is not None else 1.0
    ctx.stash = None                    # ownership passes to autograd (no copy in AccumulateGrad)
    none[0], none[1], none[2] = ghm, goff, ctx.gvm0 * ((g7[6] + g7[3]) / assumed)
    del ghm, goff
    return tuple(none)


fusion_step_vmean.register_autograd(_vmean_backward, setup_context=_vmean_setup)


# --------------------------------------------------------------------------- eager callers of the ops without autograd
class _Fast:
    """`ops.fast.decode(...)` etc.: the op's implementation without torch.library's dispatch (20-40 us of host time per call —
    more than a B = 256 decode takes on the GPU); the registered op while something is being traced.  For the ops that have
    no autograd rule (encode*, decode*, refine_centroid, postprocess, coords_to_image, loss_denominators)."""
    _NAMES = ("encode", "encode_mode", "decode", "decode_argmax", "refine_centroid", "postprocess", "coords_to_image", "loss_denominators")

    def __getattr__(self, name):
        if name not in self._NAMES:
            raise AttributeError(name)
        op = globals()[name]
        impl = op._init_fn

        def call(*a, **k):
            return op(*a, **k) if torch.compiler.is_compiling() else impl(*a, **k)

        setattr(self, name, call)
        return call


fast = _Fast()
