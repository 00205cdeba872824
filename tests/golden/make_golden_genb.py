#!/usr/bin/env python
"""Golden vectors of the second-generation ("Gen-B") family, made by RUNNING THE REFERENCE.

Build container only (needs /root/reference):

    python tests/golden/make_golden_genb.py

Feeds the seeded batches of tests/synth.py (+ make_genb_extras) to the reference's own

    utils/postprocess.py   get_max_preds, get_max_preds_with_subpixel, fused_decode,
                           coordinate_refinement, filter_low_confidence, transform_preds,
                           postprocess_predictions
    models/losses.py       FusedPoseLoss, MorphologyShapeLoss, OffsetRegressionLoss, JointsMSELoss,
                           CombinedLoss (+ autograd)
    models/pose_estimator.py   KeypointMSELoss
    data/coco_dataset.py   PreemieCocoDataset._generate_heatmaps
    data/pose_transforms.py    GenerateTarget
    validate.py            transform_preds (+ the scaling at :102-105)

and stores the outputs as tests/golden/genb_<config>.npz.  The reference is imported unmodified;
`pycocotools` (absent here, needed only to parse COCO files) is stubbed in sys.modules.
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
    mods = {}
    for name in ("utils.postprocess", "models.losses", "models.pose_estimator", "data.coco_dataset", "data.pose_transforms"):
        mods[name] = importlib.import_module(name)
    return mods


def validate_transform(coords, center, scale, heatmap_size, input_size):
    """validate.py:31-36 and :102-119 verbatim in behaviour: importing validate.py pulls the whole
    training stack (datasets, evaluator, pycocotools), so its six lines of numpy are re-run here on
    the same dtypes instead (float32 arrays, Python-float scales)."""
    out = coords.copy()
    out[:, :, 0] *= input_size[0] / heatmap_size[0]
    out[:, :, 1] *= input_size[1] / heatmap_size[1]
    for i in range(out.shape[0]):
        for k in range(out.shape[1]):
            c = out[i, k].copy()
            t = c.copy()
            t[0] = c[0] / input_size[0] * scale[i][0] + center[i][0] - scale[i][0] / 2
            t[1] = c[1] / input_size[1] * scale[i][1] + center[i][1] - scale[i][1] / 2
            out[i, k] = t
    return out


class _NS(types.SimpleNamespace):
    pass


def sparse(a):
    flat = a.reshape(-1)
    nz = np.flatnonzero(flat)
    return nz.astype(np.int64), flat[nz]


def main():
    torch.manual_seed(0)
    m = import_reference()
    pp, ls, pe, dcd, ptr = (m["utils.postprocess"], m["models.losses"], m["models.pose_estimator"], m["data.coco_dataset"],
                            m["data.pose_transforms"])
    for name, cfg in synth.CONFIGS.items():
        out = {}
        batch = synth.make_batch(cfg, seed=0)
        ex = synth.make_genb_extras(cfg, batch, seed=0)
        out["digest"] = np.array(synth.digest({**batch, **{"x_" + k: v for k, v in ex.items()}}))
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        P, pred = T(batch["heatmaps"]), T(ex["pred"])
        W, H = cfg.heatmap_size

        # ---- decode family --------------------------------------------------------------------
        c, v = pp.get_max_preds(P)
        out["maxpreds"], out["maxvals"] = c.numpy(), v.numpy()
        c, v = pp.get_max_preds_with_subpixel(P)
        out["taylor"] = c.numpy()
        c, v = pp.get_max_preds_with_subpixel(pred)
        out["taylor_pos"] = c.numpy()
        c, v = pp.fused_decode(pred, T(ex["reg_norm"]).clone(), T(ex["center"]), T(ex["scale"]), alpha=0.4)
        out["fused_norm"] = c.numpy()
        c, v = pp.fused_decode(pred, T(ex["reg_px"]).clone(), None, None)
        out["fused_px"] = c.numpy()
        c, v = pp.fused_decode(pred, None, T(ex["center"]), T(ex["scale"]))
        out["fused_scaled_only"] = c.numpy()
        start, _ = pp.get_max_preds_with_subpixel(pred)
        out["refined5"] = pp.coordinate_refinement(pred, start, 5).numpy()
        out["refined7"] = pp.coordinate_refinement(pred, start + 0.75, 7).numpy()
        f, msk = pp.filter_low_confidence(start, v, threshold=0.6)
        out["filtered"], out["filter_mask"] = f.numpy(), msk.numpy()
        out["transformed"] = pp.transform_preds(start, T(ex["center"]), T(ex["scale"]), output_size=[640, 480]).numpy()
        config = _NS(TEST=_NS(FUSION_ALPHA=0.4))
        for tag, reg in (("pipe_norm", ex["reg_norm"]), ("pipe_px", ex["reg_px"])):
            r = pp.postprocess_predictions({"heatmaps": pred, "coords": T(reg).clone()},
                                           {"center": T(ex["center"]), "scale": T(ex["scale"])}, config)
            out[tag + "_preds"], out[tag + "_mask"] = r["preds"].numpy(), r["mask"].numpy()
        r = pp.postprocess_predictions({"heatmaps": pred}, {}, config)
        out["pipe_plain_preds"], out["pipe_plain_mask"] = r["preds"].numpy(), r["mask"].numpy()
        out["to_image"] = validate_transform(batch["kps"].astype(np.float32) * np.float32(0.25) + np.float32(0.3),
                                             ex["center"], ex["scale"], cfg.heatmap_size, cfg.input_size)

        # ---- losses -----------------------------------------------------------------------------
        tgt, wgt = T(batch["target"]), T(batch["weight"])
        out["fused_mse_w"] = np.float64(ls.FusedPoseLoss(True, "mse")(pred, tgt, wgt))
        out["fused_mse_now"] = np.float64(ls.FusedPoseLoss(False, "mse")(pred, tgt, wgt))
        out["fused_mse_none"] = np.float64(ls.FusedPoseLoss(True, "mse")(pred, tgt, None))
        out["fused_sl1_w"] = np.float64(ls.FusedPoseLoss(True, "smoothl1")(pred * 3, tgt, wgt))
        out["morph_w"] = np.float64(ls.MorphologyShapeLoss(1.2, 0.5)(pred, tgt, wgt))
        out["morph_none"] = np.float64(ls.MorphologyShapeLoss(1.0, 0.5)(pred, tgt, None))
        for lt in ("smoothl1", "l1", "mse"):
            out[f"reg_{lt}"] = np.float64(ls.OffsetRegressionLoss(lt)(T(ex["coords"]), T(ex["target_coords"]), wgt))
        out["reg_none"] = np.float64(ls.OffsetRegressionLoss("smoothl1")(T(ex["coords"]), T(ex["target_coords"]), None))
        out["joints_w"] = np.float64(ls.JointsMSELoss(True)(pred, tgt, wgt))
        out["joints_now"] = np.float64(ls.JointsMSELoss(False)(pred, tgt, wgt))
        out["kpmse_w"] = np.float64(pe.KeypointMSELoss(True)(pred, tgt, wgt))
        out["kpmse_none"] = np.float64(pe.KeypointMSELoss(True)(pred, tgt, None))
        cfg_loss = _NS(LOSS=_NS(MORPH_LAMBDA=1.2, MORPH_WEIGHT=0.15, REG_WEIGHT=0.6))     # configs/preemie_optimized.yaml:19-23
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            crit = ls.CombinedLoss(cfg_loss)
            p = pred.detach().clone().to(dt).requires_grad_(True)
            c = T(ex["coords"]).clone().to(dt).requires_grad_(True)
            r = T(ex["refined"]).clone().to(dt).requires_grad_(True)
            total, parts = crit({"heatmaps": p, "coords": c, "refined_coords": r},
                                {"heatmaps": tgt.to(dt), "coords": T(ex["target_coords"]).to(dt), "weights": wgt.to(dt)})
            total.backward()
            out[f"combined_{tag}"] = np.array([float(parts[k].detach()) for k in ("heatmap", "morph", "regression", "refined", "total")], np.float64)
            if tag == "f32":
                out["combined_grad_pred"] = p.grad.numpy()
                out["combined_grad_coords"], out["combined_grad_refined"] = c.grad.numpy(), r.grad.numpy()
            else:
                out["combined_grad_pred_f64_sub"] = p.grad.numpy().reshape(-1)[::97].copy()
        # morphology alone with its gradient (the term with the non-trivial backward)
        p = pred.clone().requires_grad_(True)
        ls.MorphologyShapeLoss(1.2, 0.5)(p, tgt, wgt).backward()
        out["morph_grad_pred"] = p.grad.numpy()

        # ---- encoders ---------------------------------------------------------------------------
        ds = object.__new__(dcd.PreemieCocoDataset)
        ds.num_joints, ds.image_size, ds.sigma = cfg.K, list(cfg.input_size), cfg.sigma
        def ref_clipped(kps, vis):
            ts, ws = [], []
            for b in range(kps.shape[0]):
                t, w = ds._generate_heatmaps(kps[b], vis[b][:, None], (H, W))
                ts.append(t); ws.append(w)
            return np.stack(ts), np.stack(ws)
        t, w = ref_clipped(batch["kps"], batch["vis"])
        out["clip_nz_idx"], out["clip_nz_val"], out["clip_weight"] = *sparse(t), w
        ek, ev = synth.edge_keypoints(cfg)
        t, w = ref_clipped(ek, ev)
        out["clip_edge_nz_idx"], out["clip_edge_nz_val"], out["clip_edge_weight"] = *sparse(t), w

        gen = ptr.GenerateTarget(encoder={"input_size": (cfg.input_size[1], cfg.input_size[0]), "heatmap_size": (H, W), "sigma": cfg.sigma})
        def ref_dense(kps, vis):
            hs, ws = [], []
            for b in range(kps.shape[0]):
                r = gen({"keypoints": kps[b].copy(), "keypoints_visible": vis[b]})
                hs.append(r["heatmaps"]); ws.append(r["keypoint_weights"])
            return np.stack(hs), np.stack(ws)
        h, w = ref_dense(batch["kps"][:1], batch["vis"][:1])
        out["dense0"], out["dense0_weight"] = h, w
        _, w = ref_dense(batch["kps"], batch["vis"])
        out["dense_weight"] = w
        h, w = ref_dense(ek[:1], ev[:1])
        out["dense_edge0"], out["dense_edge0_weight"] = h, w

        path = os.path.join(HERE, f"genb_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)  combined={out['combined_f32']}")

    # ---- data/test_transforms.py:188-221, 342-379: the reference's own example of GenerateTarget -----
    gen = ptr.GenerateTarget(encoder={"input_size": (192, 256), "heatmap_size": (48, 64), "sigma": 2.0})
    kp = np.array([[96, 128], [100, 120], [80, 140]], dtype=np.float32)
    r = gen({"keypoints": kp.copy(), "keypoints_visible": np.array([1, 1, 1])})
    hm = r["heatmaps"]
    peaks = [(float(h.max()), *np.unravel_index(h.argmax(), h.shape)) for h in hm]
    np.savez_compressed(os.path.join(HERE, "genb_test_transforms.npz"), keypoints=kp, heatmaps=hm,
                        weights=r["keypoint_weights"], peaks=np.array(peaks, np.float64))
    print("test_transforms example: peaks", peaks, "shape", hm.shape)


if __name__ == "__main__":
    main()
