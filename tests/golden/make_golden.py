#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

For every config in tests/synth.py it draws the seeded batch, feeds it to the
reference's own entry points

    COCOPoseDataset._generate_target      datasets/coco_dataset.py:185
    FusionPoseLoss.forward + autograd     models/fusion_head.py:745
    HeatmapRegressionHead.decode          models/fusion_head.py:309
    PoseEstimator.inference (flip branch) models/pose_estimator.py:275
    PoseEstimator.decode_heatmaps         models/pose_estimator.py:331

and stores their outputs (inputs are regenerated from the seed by the tests;
only kps/vis and a digest of the full input set are stored).  The reference is
imported unmodified; `pycocotools` (absent here, needed only for COCO file
parsing) is stubbed in sys.modules so that datasets.coco_dataset imports.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GBCODEC_REF", "/root/reference")
sys.path.insert(0, ROOT)

from tests import synth  # noqa: E402


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference tree not found at {REF}")
    for name in ("pycocotools", "pycocotools.coco"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pycocotools.coco"].COCO = object
    sys.path.insert(0, REF)
    import importlib
    fh = importlib.import_module("models.fusion_head")
    pe = importlib.import_module("models.pose_estimator")
    ds = importlib.import_module("datasets.coco_dataset")
    return fh, pe, ds


def ref_encode(ds_mod, cfg, kps, vis):
    ds = object.__new__(ds_mod.COCOPoseDataset)
    ds.input_size = np.array(cfg.input_size)
    ds.heatmap_size = np.array(cfg.heatmap_size)
    ds.sigma = cfg.sigma
    ds.num_keypoints = cfg.K
    ts, ws = [], []
    for b in range(kps.shape[0]):
        t, w = ds._generate_target(kps[b], vis[b])
        ts.append(t); ws.append(w)
    return np.stack(ts), np.stack(ws)


class _CannedModel:
    """Runs the reference's PoseEstimator.inference with forward() replaced by
    canned head outputs (first call: un-flipped pass, second: flipped pass)."""
    def __init__(self, pe_mod, head, passes):
        self.head_type = "fusion"
        self.head = head
        self._passes = list(passes)
        self._pe = pe_mod
        self.decode_heatmaps = pe_mod.PoseEstimator.decode_heatmaps

    def forward(self, x):
        return dict(self._passes.pop(0))

    def inference(self, x, flip=True, flip_pairs=None):
        return self._pe.PoseEstimator.inference(self, x, flip=flip, flip_pairs=flip_pairs)


def sparse(a: np.ndarray):
    flat = a.reshape(-1)
    nz = np.flatnonzero(flat)
    return nz.astype(np.int64), flat[nz]


# Second set, at BASELINE.json configs[0] size (B = 32 at 64x48) and B = 8 for the two larger tile shapes.  The full
# heatmap gradient would be 6.7 MB per file, so it is stored as (a) every 7th element and (b) six float64 moments per
# tile (sum g, sum |g|, sum g^2, max |g|, sum g x, sum g y): every tile and every pixel takes part in the comparison.
LARGE = {"w32_256x192": 32, "hrformer_384x288": 8, "preemie_256": 8}
LARGE_SEED, LARGE_STRIDE = 1, 7


def tile_moments(g: np.ndarray) -> np.ndarray:
    g = g.astype(np.float64)
    H, W = g.shape[-2:]
    xs, ys = np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64)
    return np.stack([g.sum((2, 3)), np.abs(g).sum((2, 3)), (g * g).sum((2, 3)), np.abs(g).max((2, 3)),
                     (g * xs[None, None, None, :]).sum((2, 3)), (g * ys[None, None, :, None]).sum((2, 3))], axis=-1)


def main(large: bool = False):
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    fh, pe, dsm = import_reference()
    from oracle import heatmap_codec as oc

    for name, cfg in synth.CONFIGS.items():
        out = {}
        if large:
            import dataclasses
            cfg = dataclasses.replace(cfg, B=LARGE[name])
            batch = synth.make_batch(cfg, seed=LARGE_SEED, B=cfg.B)
        else:
            batch = synth.make_batch(cfg, seed=0)
        out["digest"] = np.array(synth.digest(batch))
        out["kps"], out["vis"] = batch["kps"], batch["vis"]

        # ---- encode (random + edge set) -------------------------------------
        t_ref, w_ref = ref_encode(dsm, cfg, batch["kps"], batch["vis"])
        assert np.array_equal(t_ref, batch["target"]) and np.array_equal(w_ref, batch["weight"]), \
            "oracle encode differs from the reference on the seeded batch"
        out["enc_weight"] = w_ref
        out["enc_nz_idx"], out["enc_nz_val"] = sparse(t_ref)
        ek, ev = synth.edge_keypoints(cfg)
        et, ew = ref_encode(dsm, cfg, ek, ev)
        out["edge_kps"], out["edge_vis"], out["edge_weight"] = ek, ev, ew
        out["edge_nz_idx"], out["edge_nz_val"] = sparse(et)

        # ---- loss forward + autograd ----------------------------------------
        T = lambda k: torch.from_numpy(batch[k])
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            loss_fn = fh.FusionPoseLoss(target_sigma=cfg.sigma)
            h = T("heatmaps").to(dt).requires_grad_(True)
            o = T("offsets").to(dt).requires_grad_(True)
            v = T("variances").to(dt).requires_grad_(True)
            losses = loss_fn({"heatmaps": h, "offsets": o, "variances": v},
                             torch.from_numpy(t_ref).to(dt), torch.from_numpy(w_ref).to(dt),
                             T("kps").to(dt), input_size=cfg.input_size, heatmap_size=(cfg.H, cfg.W))
            losses["total_loss"].backward()
            out[f"loss_{tag}"] = np.array([float(losses[k].detach()) for k in oc.LOSS_KEYS], np.float64)
            if tag == "f32" and large:
                out["grad_hm_sub"] = h.grad.numpy().reshape(-1)[::LARGE_STRIDE].copy()
                out["grad_hm_moments"] = tile_moments(h.grad.numpy())
                out["grad_off_idx"], out["grad_off_val"] = sparse(o.grad.numpy())
                out["grad_var_tile"] = v.grad.numpy()[:, :, 0, 0].copy()
            elif tag == "f32":
                out["grad_hm"] = h.grad.numpy()
                out["grad_off_idx"], out["grad_off_val"] = sparse(o.grad.numpy())
                out["grad_var_tile"] = v.grad.numpy()[:, :, 0, 0].copy()
                assert np.all(v.grad.numpy() == v.grad.numpy()[:, :, :1, :1]), "grad_var not uniform per tile"
            else:
                out["grad_hm_f64_absmax"] = np.abs(h.grad.numpy()).max(axis=(2, 3))
                # store the f64 gradient at a strided subset: an error bar for the f32 reference itself
                out["grad_hm_f64_sub"] = h.grad.numpy().reshape(-1)[::97].copy()
                out["grad_off_f64_idx"], out["grad_off_f64_val"] = sparse(o.grad.numpy())
                out["grad_var_f64_tile"] = v.grad.numpy()[:, :, 0, 0].copy()

        # ---- decode ------------------------------------------------------------
        head = fh.HeatmapRegressionHead(in_channels=8, num_keypoints=cfg.K, hidden_dim=8)
        head.eval()
        with torch.no_grad():
            fw = torch.sigmoid(head.fusion_weight)
            outputs = {"heatmaps": T("heatmaps"), "offsets": T("offsets"), "variances": T("variances"),
                       "fusion_weight": fw}
            c, s = head.decode(outputs, apply_offset=True)
            out["dec_coords"], out["dec_scores"] = c.numpy(), s.numpy()
            c, s = head.decode(outputs, apply_offset=False)
            out["dec_coords_nooff"] = c.numpy()
            g, _ = fh.SoftArgmax2D()(T("heatmaps"))
            out["dec_softargmax"] = g.numpy()
            # the model sees the flipped image, so its raw output is what inference() flips back:
            raw_flipped_pass = {"heatmaps": T("heatmaps_flip"), "offsets": T("offsets"),
                                "variances": T("variances"), "fusion_weight": fw}
            model = _CannedModel(pe, head, [outputs, raw_flipped_pass])
            pairs = [p for p in oc.COCO_FLIP_PAIRS if p[0] < cfg.K and p[1] < cfg.K]
            c, s = model.inference(torch.zeros(cfg.B, 3, 4, 4), flip=True, flip_pairs=pairs)
            out["flip_coords"], out["flip_scores"] = c.numpy(), s.numpy()
            c, s = pe.PoseEstimator.decode_heatmaps(T("heatmaps"), shift=True)
            out["argmax_coords"], out["argmax_vals"] = c.numpy(), s.numpy()
            c, _ = pe.PoseEstimator.decode_heatmaps(T("heatmaps"), shift=False)
            out["argmax_idx"] = (c[..., 1].long() * cfg.W + c[..., 0].long()).numpy()
        out["alpha_param"] = np.float32(head.subpixel_refine.alpha.item())
        out["fusion_weight"] = np.float32(fw.item())

        path = os.path.join(HERE, f"{name}_large.npz" if large else f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)  total_loss={out['loss_f32'][-1]:.6f}")


if __name__ == "__main__":
    main(large="--large" in sys.argv)
