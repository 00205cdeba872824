"""Load the committed golden vectors (made by tests/golden/make_golden.py from the
reference itself) together with the regenerated seeded inputs."""
from __future__ import annotations

import os
import warnings
from functools import lru_cache

import numpy as np

from tests import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dense(idx, val, shape, dtype=np.float32):
    a = np.zeros(int(np.prod(shape)), dtype)
    a[idx] = val
    return a.reshape(shape)


@lru_cache(maxsize=None)
def load(name: str):
    cfg = synth.CONFIGS[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    batch = synth.make_batch(cfg, seed=0)
    if synth.digest(batch) != str(g["digest"]):
        # numpy's SIMD exp may differ in the last bit between hosts; tolerances absorb that.
        warnings.warn(f"{name}: regenerated inputs differ bitwise from the ones the goldens were made with")
    assert np.array_equal(batch["kps"], g["kps"]) and np.array_equal(batch["vis"], g["vis"]), \
        "seeded keypoints drifted: the generator changed, regenerate tests/golden"
    B, K, H, W = cfg.B, cfg.K, cfg.H, cfg.W
    g["target"] = dense(g["enc_nz_idx"], g["enc_nz_val"], (B, K, H, W))
    eb = g["edge_kps"].shape[0]
    g["edge_target"] = dense(g["edge_nz_idx"], g["edge_nz_val"], (eb, K, H, W))
    g["grad_off"] = dense(g["grad_off_idx"], g["grad_off_val"], (B, K, 2, H, W))
    g["grad_off_f64"] = dense(g["grad_off_f64_idx"], g["grad_off_f64_val"], (B, K, 2, H, W), np.float64)
    return cfg, batch, g


@lru_cache(maxsize=None)
def load_large(name: str):
    """The second golden set (tests/golden/make_golden.py --large): BASELINE configs[0] size for the 64x48 shape (B = 32),
    B = 8 for the two larger tile shapes, seed 1.  The heatmap gradient is stored as every 7th element (`grad_hm_sub`)
    plus six float64 moments per tile (`grad_hm_moments`: sum g, sum |g|, sum g^2, max |g|, sum g x, sum g y)."""
    import dataclasses
    from tests.golden.make_golden import LARGE, LARGE_SEED
    cfg = dataclasses.replace(synth.CONFIGS[name], B=LARGE[name])
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}_large.npz")))
    batch = synth.make_batch(cfg, seed=LARGE_SEED, B=cfg.B)
    assert np.array_equal(batch["kps"], g["kps"]) and np.array_equal(batch["vis"], g["vis"]), \
        "seeded keypoints drifted: the generator changed, regenerate tests/golden (--large)"
    B, K, H, W = cfg.B, cfg.K, cfg.H, cfg.W
    g["target"] = dense(g["enc_nz_idx"], g["enc_nz_val"], (B, K, H, W))
    g["grad_off"] = dense(g["grad_off_idx"], g["grad_off_val"], (B, K, 2, H, W))
    g["grad_off_f64"] = dense(g["grad_off_f64_idx"], g["grad_off_f64_val"], (B, K, 2, H, W), np.float64)
    return cfg, batch, g
