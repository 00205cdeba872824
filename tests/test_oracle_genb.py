"""CPU: the Gen-B restatement (oracle/genb.py) against the golden vectors made by running the
reference (tests/golden/make_golden_genb.py), and against the peaks the reference's own
data/test_transforms.py prints for its GenerateTarget example."""
import numpy as np
import pytest
import torch

from oracle import genb
from tests import goldens_genb, synth

NAMES = list(synth.CONFIGS)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("name", NAMES)
def test_decode_family(name):
    cfg, batch, ex, g = goldens_genb.load(name)
    P, pred = T(batch["heatmaps"]), T(ex["pred"])
    c, v = genb.get_max_preds(P)
    assert np.array_equal(c.numpy(), g["maxpreds"]) and np.array_equal(v.numpy(), g["maxvals"])
    assert np.array_equal(genb.get_max_preds_with_subpixel(P)[0].numpy(), g["taylor"])
    start, v = genb.get_max_preds_with_subpixel(pred)
    assert np.array_equal(start.numpy(), g["taylor_pos"])
    c, _ = genb.fused_decode(pred, T(ex["reg_norm"]), T(ex["center"]), T(ex["scale"]), alpha=0.4)
    np.testing.assert_allclose(c.numpy(), g["fused_norm"], rtol=0, atol=1e-5)
    c, _ = genb.fused_decode(pred, T(ex["reg_px"]), None, None)
    np.testing.assert_allclose(c.numpy(), g["fused_px"], rtol=0, atol=1e-5)
    c, _ = genb.fused_decode(pred, None, T(ex["center"]), T(ex["scale"]))
    assert np.array_equal(c.numpy(), g["fused_scaled_only"])
    np.testing.assert_allclose(genb.coordinate_refinement(pred, start, 5).numpy(), g["refined5"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(genb.coordinate_refinement(pred, start + 0.75, 7).numpy(), g["refined7"], rtol=0, atol=1e-5)
    f, m = genb.filter_low_confidence(start, v, 0.6)
    assert np.array_equal(f.numpy(), g["filtered"]) and np.array_equal(m.numpy(), g["filter_mask"])
    assert np.array_equal(genb.transform_preds(start, T(ex["center"]), T(ex["scale"])).numpy(), g["transformed"])
    for tag, reg in (("pipe_norm", ex["reg_norm"]), ("pipe_px", ex["reg_px"])):
        r = genb.postprocess_predictions(pred, T(reg), T(ex["center"]), T(ex["scale"]), alpha=0.4)
        np.testing.assert_allclose(r["preds"].numpy(), g[tag + "_preds"], rtol=0, atol=2e-4)
        assert np.array_equal(r["mask"].numpy(), g[tag + "_mask"])
    r = genb.postprocess_predictions(pred)
    np.testing.assert_allclose(r["preds"].numpy(), g["pipe_plain_preds"], rtol=0, atol=1e-5)
    got = genb.heatmap_to_image(batch["kps"].astype(np.float32) * np.float32(0.25) + np.float32(0.3), ex["center"], ex["scale"],
                                cfg.heatmap_size, cfg.input_size)
    assert np.array_equal(got, g["to_image"])


@pytest.mark.parametrize("name", NAMES)
def test_losses(name):
    cfg, batch, ex, g = goldens_genb.load(name)
    pred, tgt, wgt = T(ex["pred"]), T(batch["target"]), T(batch["weight"])
    close = lambda got, key: np.testing.assert_allclose(float(got), float(g[key]), rtol=2e-6, err_msg=key)
    close(genb.fused_pose_loss(pred, tgt, wgt, True, "mse"), "fused_mse_w")
    close(genb.fused_pose_loss(pred, tgt, wgt, False, "mse"), "fused_mse_now")
    close(genb.fused_pose_loss(pred, tgt, None, True, "mse"), "fused_mse_none")
    close(genb.fused_pose_loss(pred * 3, tgt, wgt, True, "smoothl1"), "fused_sl1_w")
    close(genb.morphology_shape_loss(pred, tgt, wgt, 1.2, 0.5), "morph_w")
    close(genb.morphology_shape_loss(pred, tgt, None, 1.0, 0.5), "morph_none")
    for lt in ("smoothl1", "l1", "mse"):
        close(genb.offset_regression_loss(T(ex["coords"]), T(ex["target_coords"]), wgt, lt), f"reg_{lt}")
    close(genb.offset_regression_loss(T(ex["coords"]), T(ex["target_coords"]), None), "reg_none")
    close(genb.joints_mse_loss(pred, tgt, wgt, True), "joints_w")
    close(genb.joints_mse_loss(pred, tgt, wgt, False), "joints_now")
    close(genb.keypoint_mse_loss(pred, tgt, wgt, True), "kpmse_w")
    close(genb.keypoint_mse_loss(pred, tgt, None, True), "kpmse_none")
    p = pred.clone().requires_grad_(True)
    c = T(ex["coords"]).clone().requires_grad_(True)
    r = T(ex["refined"]).clone().requires_grad_(True)
    total, parts = genb.combined_loss({"heatmaps": p, "coords": c, "refined_coords": r},
                                      {"heatmaps": tgt, "coords": T(ex["target_coords"]), "weights": wgt},
                                      morph_lambda=1.2, morph_weight=0.15, reg_weight=0.6)
    total.backward()
    got = np.array([float(parts[k].detach()) for k in genb.COMBINED_KEYS])
    np.testing.assert_allclose(got, g["combined_f32"], rtol=2e-6)
    np.testing.assert_allclose(p.grad.numpy(), g["combined_grad_pred"], rtol=1e-5, atol=1e-6 * np.abs(g["combined_grad_pred"]).max())
    np.testing.assert_allclose(c.grad.numpy(), g["combined_grad_coords"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(r.grad.numpy(), g["combined_grad_refined"], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("name", NAMES)
def test_encoders(name):
    cfg, batch, ex, g = goldens_genb.load(name)
    W, H = cfg.heatmap_size
    t, w = genb.encode_patch_clipped(batch["kps"], batch["vis"], (H, W), cfg.input_size, cfg.sigma)
    assert np.array_equal(t, g["clip_target"]) and np.array_equal(w, g["clip_weight"])
    ek, ev = synth.edge_keypoints(cfg)
    t, w = genb.encode_patch_clipped(ek, ev, (H, W), cfg.input_size, cfg.sigma)
    assert np.array_equal(t, g["clip_edge_target"]) and np.array_equal(w, g["clip_edge_weight"])
    isz = (cfg.input_size[1], cfg.input_size[0])
    h, w = genb.encode_dense(batch["kps"], batch["vis"], (H, W), isz, cfg.sigma)
    assert np.array_equal(h[:1], g["dense0"]) and np.array_equal(w, g["dense_weight"])
    h, w = genb.encode_dense(ek[:1], ev[:1], (H, W), isz, cfg.sigma)
    assert np.array_equal(h, g["dense_edge0"]) and np.array_equal(w, g["dense_edge0_weight"])


def test_generate_target_example_of_the_reference():
    """data/test_transforms.py:342-379 prints, for keypoints (96,128), (100,120), (80,140) and the codec
    input (192,256) / heatmap (48,64) / sigma 2, the peak value and position of every heatmap.  Running
    that example on the reference gives peak 1.0 at (row, col) = (32,24), (30,25), (35,20) on a
    48-row x 64-column map (the (H,W)/(W,H) swap of that class) — the only printed known answers in
    the reference for this path."""
    g = goldens_genb.load_test_transforms()
    h, w = genb.encode_dense(g["keypoints"][None], np.ones((1, 3), np.float32), (48, 64), (192, 256), 2.0)
    assert h.shape == (1, 3, 48, 64)
    assert np.array_equal(h[0], g["heatmaps"]) and np.array_equal(w[0], g["weights"])
    peaks = [(float(t.max()), *np.unravel_index(t.argmax(), t.shape)) for t in h[0]]
    assert peaks == [(1.0, 32, 24), (1.0, 30, 25), (1.0, 35, 20)]
    assert np.array_equal(np.array(peaks, np.float64), g["peaks"])
