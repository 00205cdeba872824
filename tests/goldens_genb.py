"""Load the committed Gen-B golden vectors (tests/golden/make_golden_genb.py ran the reference)."""
from __future__ import annotations

import os
import warnings
from functools import lru_cache

import numpy as np

from tests import synth
from tests.goldens import GOLDEN_DIR, dense


@lru_cache(maxsize=None)
def load(name: str):
    cfg = synth.CONFIGS[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"genb_{name}.npz")))
    batch = synth.make_batch(cfg, seed=0)
    ex = synth.make_genb_extras(cfg, batch, seed=0)
    if synth.digest({**batch, **{"x_" + k: v for k, v in ex.items()}}) != str(g["digest"]):
        warnings.warn(f"genb {name}: regenerated inputs differ bitwise from the ones the goldens were made with")
    B, K, H, W = cfg.B, cfg.K, cfg.H, cfg.W
    g["clip_target"] = dense(g["clip_nz_idx"], g["clip_nz_val"], (B, K, H, W))
    ek, _ = synth.edge_keypoints(cfg)
    g["clip_edge_target"] = dense(g["clip_edge_nz_idx"], g["clip_edge_nz_val"], (ek.shape[0], K, H, W))
    return cfg, batch, ex, g


def load_test_transforms():
    return dict(np.load(os.path.join(GOLDEN_DIR, "genb_test_transforms.npz")))
