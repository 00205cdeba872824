"""ctypes binding of libgbcodec.so (include/gbcodec.h).

There is no fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional, Sequence, Tuple

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libgbcodec.so")

MAX_K = 64
MAX_PAIRS = 64
MAX_PARTNERS = 4
MAX_TILE = 32768
PEER_HANDLE_BYTES = 64
MAX_PEERS = 16
ABI_VERSION = 3

DECODE_REFINE = 1
DECODE_APPLY_OFFSET = 2
DECODE_FUSION_WEIGHT_RAW = 4
ARGMAX_PLAIN, ARGMAX_QUARTER, ARGMAX_TAYLOR = 0, 1, 2
ENCODE_PATCH, ENCODE_PATCH_CLIPPED, ENCODE_DENSE = 0, 1, 2
CRIT_MSE, CRIT_SMOOTHL1, CRIT_L1, CRIT_MSE_WEIGHTED = 0, 1, 2, 3
TERM_HEATMAP, TERM_MORPH, TERM_REGRESSION, TERM_REFINED = 1, 2, 4, 8

EXPORTS = (
    "gbcodec_abi_version", "gbcodec_status_string", "gbcodec_last_error", "gbcodec_launch_count",
    "gbcodec_encode_f32", "gbcodec_decode_f32", "gbcodec_decode_argmax_f32", "gbcodec_refine_centroid_f32",
    "gbcodec_loss_workspace_bytes", "gbcodec_loss_denominators_f32",
    "gbcodec_fusion_loss_f32", "gbcodec_fusion_step_f32", "gbcodec_fusion_loss_backward_f32",
    "gbcodec_profile_loss_kernel",
    "gbcodec_encode_mode_f32", "gbcodec_postprocess_f32", "gbcodec_coords_to_image_f32",
    "gbcodec_combined_workspace_bytes", "gbcodec_combined_loss_f32", "gbcodec_combined_loss_backward_f32",
    "gbcodec_peer_create", "gbcodec_peer_connect", "gbcodec_peer_status", "gbcodec_peer_destroy", "gbcodec_peer_set_timeout",
    "gbcodec_peer_denominators_f32", "gbcodec_peer_collect_losses_f32",
    "gbcodec_softplus_mean_f32", "gbcodec_softplus_mean_backward_f32",
    "gbcodec_fusion_step_sharded_f32", "gbcodec_heatmap_step_f32",
    "gbcodec_fusion_step_f16", "gbcodec_fusion_loss_backward_f16", "gbcodec_fusion_step_vmean_f32",
    "gbcodec_combined_loss_f16", "gbcodec_combined_loss_backward_f16",
)


class LossDesc(C.Structure):
    """struct gbcodec_loss_desc"""
    _fields_ = [
        ("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("in_w", C.c_float), ("in_h", C.c_float),
        ("lambdas", C.c_float * 6),
        ("target_sigma", C.c_double), ("encode_sigma", C.c_double),
        ("use_target_weight", C.c_int32), ("n_pairs", C.c_int32),
        ("pairs", (C.c_int32 * 2) * MAX_PAIRS),
    ]


class PostprocessDesc(C.Structure):
    """struct gbcodec_postprocess_desc"""
    _fields_ = [
        ("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("argmax_mode", C.c_int32), ("scale_to_image", C.c_int32), ("image_size", C.c_float),
        ("refine_window", C.c_int32), ("filter", C.c_int32), ("threshold", C.c_float),
        ("transform", C.c_int32), ("input_w", C.c_float), ("input_h", C.c_float),
    ]


class CombinedDesc(C.Structure):
    """struct gbcodec_combined_desc"""
    _fields_ = [
        ("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("norm_batch", C.c_int32), ("terms", C.c_uint32),
        ("heatmap_criterion", C.c_int32), ("coord_criterion", C.c_int32), ("use_target_weight", C.c_int32),
        ("heatmap_scale", C.c_float), ("lambda_variance", C.c_float), ("lambda_mean", C.c_float),
        ("w_heatmap", C.c_float), ("w_morph", C.c_float), ("w_reg", C.c_float),
    ]


def make_loss_desc(B: int, K: int, H: int, W: int, in_w: float, in_h: float, lambdas: Sequence[float],
                   target_sigma: float, encode_sigma: float, use_target_weight: bool,
                   pairs: Sequence[Tuple[int, int]]) -> LossDesc:
    if len(lambdas) != 6:
        raise ValueError("lambdas must hold six weights")
    if len(pairs) > MAX_PAIRS:
        raise ValueError(f"at most {MAX_PAIRS} limb pairs")
    d = LossDesc()
    d.B, d.K, d.H, d.W = B, K, H, W
    d.in_w, d.in_h = float(in_w), float(in_h)
    for i, v in enumerate(lambdas):
        d.lambdas[i] = float(v)
    d.target_sigma, d.encode_sigma = float(target_sigma), float(encode_sigma)
    d.use_target_weight = int(bool(use_target_weight))
    d.n_pairs = len(pairs)
    for i, (a, b) in enumerate(pairs):
        d.pairs[i][0], d.pairs[i][1] = int(a), int(b)
    return d


_lib = None
_lock = threading.Lock()
_P = C.c_void_p


def _declare(lib):
    f32p = _P
    lib.gbcodec_abi_version.restype = C.c_int
    lib.gbcodec_launch_count.restype = C.c_ulonglong
    lib.gbcodec_launch_count.argtypes = []
    lib.gbcodec_status_string.restype = C.c_char_p
    lib.gbcodec_status_string.argtypes = [C.c_int]
    lib.gbcodec_last_error.restype = C.c_char_p
    lib.gbcodec_encode_f32.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_float, C.c_float, C.c_double, _P]
    lib.gbcodec_decode_f32.argtypes = [f32p, f32p, _P, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_uint, f32p, f32p, _P, _P]
    lib.gbcodec_decode_argmax_f32.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, _P, _P]
    lib.gbcodec_refine_centroid_f32.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, _P]
    lib.gbcodec_loss_workspace_bytes.restype = C.c_size_t
    lib.gbcodec_loss_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.gbcodec_loss_denominators_f32.argtypes = [C.POINTER(LossDesc), f32p, f32p, C.c_int, f32p, _P, C.c_size_t, _P]
    loss_common = [C.POINTER(LossDesc), f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p]
    lib.gbcodec_fusion_loss_f32.argtypes = loss_common + [f32p, f32p, f32p, f32p, _P, C.c_size_t, _P]
    lib.gbcodec_fusion_step_f32.argtypes = loss_common + [f32p, f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_uint,
                                                          f32p, f32p, _P, C.c_size_t, _P]
    lib.gbcodec_fusion_loss_backward_f32.argtypes = loss_common + [f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int, _P, C.c_size_t, _P]
    lib.gbcodec_profile_loss_kernel.argtypes = [_P, _P]
    lib.gbcodec_heatmap_step_f32.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                             C.c_double, C.c_int, C.c_int, f32p, f32p, f32p, C.c_int, f32p, f32p, _P,
                                             _P, C.c_size_t, _P]
    lib.gbcodec_fusion_step_f16.argtypes = [C.POINTER(LossDesc), _P, _P, _P, f32p, f32p, f32p, f32p, f32p, f32p, _P, _P, _P,
                                            f32p, f32p, C.c_int, C.c_uint, f32p, f32p, _P, C.c_size_t, _P]
    lib.gbcodec_fusion_loss_backward_f16.argtypes = [C.POINTER(LossDesc), _P, _P, _P, f32p, f32p, f32p, f32p, f32p, C.c_int, f32p,
                                                     _P, _P, _P, f32p, C.c_int, C.c_int, _P, C.c_size_t, _P]
    lib.gbcodec_fusion_step_vmean_f32.argtypes = loss_common + [f32p, f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_uint,
                                                                f32p, f32p, _P, C.c_size_t, _P]
    lib.gbcodec_peer_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_char_p]
    lib.gbcodec_peer_connect.argtypes = [_P, C.c_char_p]
    lib.gbcodec_peer_status.argtypes = [_P, C.POINTER(C.c_int)]
    lib.gbcodec_peer_set_timeout.argtypes = [_P, C.c_double]
    lib.gbcodec_peer_destroy.argtypes = [_P]
    lib.gbcodec_fusion_step_sharded_f32.argtypes = [C.POINTER(LossDesc), f32p, f32p, f32p, f32p, f32p, f32p, f32p,
                                                    f32p, f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_uint,
                                                    f32p, f32p, f32p, f32p, C.c_int, _P, C.c_size_t, _P, _P]
    lib.gbcodec_peer_denominators_f32.argtypes = [C.POINTER(LossDesc), f32p, f32p, C.c_int, f32p, _P, C.c_size_t, _P, _P]
    lib.gbcodec_peer_collect_losses_f32.argtypes = [_P, C.c_int, f32p, _P]
    lib.gbcodec_softplus_mean_f32.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, _P]
    lib.gbcodec_softplus_mean_backward_f32.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, _P]
    lib.gbcodec_encode_mode_f32.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_float, C.c_float, C.c_double, C.c_int, _P]
    lib.gbcodec_postprocess_f32.argtypes = [C.POINTER(PostprocessDesc), f32p, f32p, f32p, f32p, f32p, f32p, f32p, _P, _P]
    lib.gbcodec_coords_to_image_f32.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                                f32p, _P]
    lib.gbcodec_combined_workspace_bytes.restype = C.c_size_t
    lib.gbcodec_combined_workspace_bytes.argtypes = [C.c_int, C.c_int]
    comb = [C.POINTER(CombinedDesc), f32p, f32p, f32p, f32p, f32p, f32p, f32p]
    lib.gbcodec_combined_loss_f32.argtypes = comb + [f32p, f32p, f32p, f32p, _P, C.c_size_t, _P]
    lib.gbcodec_combined_loss_backward_f32.argtypes = comb + [f32p, f32p, f32p, f32p, _P, C.c_size_t, _P]
    # float16 predictions: d_pred_f16 and d_grad_pred_f16 are void*
    comb16 = [C.POINTER(CombinedDesc), _P, f32p, f32p, f32p, f32p, f32p, f32p]
    lib.gbcodec_combined_loss_f16.argtypes = comb16 + [f32p, _P, f32p, f32p, _P, C.c_size_t, _P]
    lib.gbcodec_combined_loss_backward_f16.argtypes = comb16 + [f32p, _P, f32p, f32p, _P, C.c_size_t, _P]
    for name in EXPORTS:
        getattr(lib, name)          # AttributeError here = the library does not export what the header declares


def lib():
    """The loaded library.  Raises if it has not been built — there is no CPU path."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with "
                        "`python -m infantposeestimation_gaussianbias_b200.build` (needs nvcc). "
                        "This package has no CPU or PyTorch fallback.")
                l = C.CDLL(LIB_PATH)
                _declare(l)
                got = l.gbcodec_abi_version()
                if got != ABI_VERSION:
                    raise RuntimeError(f"libgbcodec.so ABI {got}, binding expects {ABI_VERSION}: rebuild")
                _lib = l
    return _lib


class GbcodecError(RuntimeError):
    def __init__(self, status: int, where: str):
        l = lib()
        self.status = status
        what = l.gbcodec_status_string(status).decode()
        detail = l.gbcodec_last_error().decode()
        super().__init__(f"gbcodec {where}: {what} ({status}): {detail}")


def check(status: int, where: str) -> None:
    if status != 0:
        raise GbcodecError(status, where)
