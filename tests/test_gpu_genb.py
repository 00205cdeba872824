"""GPU parity of the second-generation ("Gen-B") family — utils/postprocess.py, models/losses.py,
KeypointMSELoss, the two alternate encoders and the coordinate transform — through the torch ops ->
ctypes -> C ABI of libgbcodec.so, against the golden vectors made by running the reference and
against the CPU oracle on larger seeded batches.  Tolerances are BASELINE.json's:

  integer peak indices, encoder weights and patch support   bit-exact
  encoded heatmaps                        1e-6 relative (fp32 exp)
  loss values and gradients               1e-5 relative
  decoded coordinates                     1e-4 px (heatmap pixels; image-space outputs: 1e-4 px + 2 ulp)
"""
import types

import numpy as np
import pytest
import torch

from oracle import genb
from tests import goldens_genb, synth

pytestmark = pytest.mark.gpu

NAMES = list(synth.CONFIGS)
LOSS_RTOL, COORD_ATOL = 1e-5, 1e-4


@pytest.fixture(scope="module")
def pkg():
    import infantposeestimation_gaussianbias_b200 as p
    p.load()
    return p


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def same(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    bad = np.argwhere(got != want)
    assert len(bad) == 0, f"{what}: {len(bad)} of {got.size} differ; first at {tuple(bad[0])}: got {got[tuple(bad[0])]!r}, want {want[tuple(bad[0])]!r}"


def maxnorm_close(got, want, rel=LOSS_RTOL, what=""):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = np.abs(want).max()
    err = np.abs(got - want).max()
    assert err <= rel * scale + 1e-30, f"{what}: max err {err:.3e} vs scale {scale:.3e} ({err / max(scale, 1e-300):.2e})"


def image_close(got, want, what=""):
    np.testing.assert_allclose(got, want, rtol=2.5e-7, atol=COORD_ATOL, err_msg=what)


# ------------------------------------------------------------------ decode family
@pytest.mark.parametrize("name", NAMES)
def test_postprocess_functions_against_golden(pkg, name):
    pp = pkg.postprocess
    cfg, batch, ex, g = goldens_genb.load(name)
    P, pred = dev(batch["heatmaps"]), dev(ex["pred"])
    c, v = pp.get_max_preds(P)
    assert np.array_equal(c.cpu().numpy(), g["maxpreds"]) and np.array_equal(v.cpu().numpy(), g["maxvals"])
    c, v = pp.get_max_preds_with_subpixel(P)
    assert np.array_equal(c.cpu().numpy(), g["taylor"]), "Taylor step is IEEE-identical arithmetic"
    start, v = pp.get_max_preds_with_subpixel(pred)
    assert np.array_equal(start.cpu().numpy(), g["taylor_pos"])
    c, _ = pp.fused_decode(pred, dev(ex["reg_norm"]), dev(ex["center"]), dev(ex["scale"]), alpha=0.4)
    image_close(c.cpu().numpy(), g["fused_norm"], "fused_decode, normalised regression branch")
    c, _ = pp.fused_decode(pred, dev(ex["reg_px"]), None, None)
    image_close(c.cpu().numpy(), g["fused_px"], "fused_decode, pixel regression branch")
    c, _ = pp.fused_decode(pred, None, dev(ex["center"]), dev(ex["scale"]))
    assert np.array_equal(c.cpu().numpy(), g["fused_scaled_only"])
    np.testing.assert_allclose(pp.coordinate_refinement(pred, start, 5).cpu().numpy(), g["refined5"], rtol=0, atol=COORD_ATOL)
    np.testing.assert_allclose(pp.coordinate_refinement(pred, start + 0.75, 7).cpu().numpy(), g["refined7"], rtol=0, atol=COORD_ATOL)
    f, m = pp.filter_low_confidence(start, v, threshold=0.6)
    assert np.array_equal(f.cpu().numpy(), g["filtered"]) and np.array_equal(m.cpu().numpy(), g["filter_mask"])
    image_close(pp.transform_preds(start, dev(ex["center"]), dev(ex["scale"]), [640, 480]).cpu().numpy(), g["transformed"])
    got = pp.heatmap_to_image(dev(batch["kps"].astype(np.float32) * np.float32(0.25) + np.float32(0.3)), dev(ex["center"]),
                              dev(ex["scale"]), cfg.heatmap_size, cfg.input_size)
    assert np.array_equal(got.cpu().numpy(), g["to_image"]), "coordinate transform follows the reference's float32 operation order"


@pytest.mark.parametrize("name", NAMES)
def test_postprocess_pipeline_against_golden(pkg, name):
    """postprocess_predictions (utils/postprocess.py:296-340) as ONE kernel."""
    pp = pkg.postprocess
    cfg, batch, ex, g = goldens_genb.load(name)
    pred = dev(ex["pred"])
    config = types.SimpleNamespace(TEST=types.SimpleNamespace(FUSION_ALPHA=0.4))
    for tag, reg in (("pipe_norm", ex["reg_norm"]), ("pipe_px", ex["reg_px"])):
        r = pp.postprocess_predictions({"heatmaps": pred, "coords": dev(reg)}, {"center": dev(ex["center"]), "scale": dev(ex["scale"])}, config)
        assert np.array_equal(r["mask"].cpu().numpy(), g[tag + "_mask"])
        image_close(r["preds"].cpu().numpy(), g[tag + "_preds"], tag)
    r = pp.postprocess_predictions({"heatmaps": pred}, {}, config)
    assert np.array_equal(r["mask"].cpu().numpy(), g["pipe_plain_mask"])
    np.testing.assert_allclose(r["preds"].cpu().numpy(), g["pipe_plain_preds"], rtol=0, atol=COORD_ATOL)


def test_postprocess_larger_batch_vs_oracle(pkg):
    pp = pkg.postprocess
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=21, B=48)
    ex = synth.make_genb_extras(cfg, batch, seed=21)
    pred = ex["pred"]
    config = types.SimpleNamespace(TEST=types.SimpleNamespace())
    # the reference's coordinate_refinement raises (torch.arange with a negative upper bound) when a fused
    # coordinate is below -window/2; keep the regression branch non-negative so that the oracle runs
    reg = np.abs(ex["reg_px"])
    r = pp.postprocess_predictions({"heatmaps": dev(pred), "coords": dev(reg)}, {"center": dev(ex["center"]), "scale": dev(ex["scale"])}, config)
    want = genb.postprocess_predictions(T(pred), T(reg), T(ex["center"]), T(ex["scale"]))
    assert np.array_equal(r["mask"].cpu().numpy(), want["mask"].numpy()) and np.array_equal(r["maxvals"].cpu().numpy(), want["maxvals"].numpy())
    image_close(r["preds"].cpu().numpy(), want["preds"].numpy())
    c, _ = pp.get_max_preds_with_subpixel(dev(batch["heatmaps"]))
    assert np.array_equal(c.cpu().numpy(), genb.get_max_preds_with_subpixel(T(batch["heatmaps"]))[0].numpy())


# ------------------------------------------------------------------ losses
@pytest.mark.parametrize("name", NAMES)
def test_losses_against_golden(pkg, name):
    L = pkg.losses
    cfg, batch, ex, g = goldens_genb.load(name)
    pred, tgt, wgt = dev(ex["pred"]), dev(batch["target"]), dev(batch["weight"])
    close = lambda got, key: np.testing.assert_allclose(float(got), float(g[key]), rtol=LOSS_RTOL, err_msg=key)
    close(L.FusedPoseLoss(True, "mse")(pred, tgt, wgt), "fused_mse_w")
    close(L.FusedPoseLoss(False, "mse")(pred, tgt, wgt), "fused_mse_now")
    close(L.FusedPoseLoss(True, "mse")(pred, tgt, None), "fused_mse_none")
    close(L.FusedPoseLoss(True, "smoothl1")(pred * 3, tgt, wgt), "fused_sl1_w")
    close(L.MorphologyShapeLoss(1.2, 0.5)(pred, tgt, wgt), "morph_w")
    close(L.MorphologyShapeLoss(1.0, 0.5)(pred, tgt, None), "morph_none")
    for lt in ("smoothl1", "l1", "mse"):
        close(L.OffsetRegressionLoss(lt)(dev(ex["coords"]), dev(ex["target_coords"]), wgt), f"reg_{lt}")
    close(L.OffsetRegressionLoss("smoothl1")(dev(ex["coords"]), dev(ex["target_coords"]), None), "reg_none")
    close(L.JointsMSELoss(True)(pred, tgt, wgt), "joints_w")
    close(L.JointsMSELoss(False)(pred, tgt, wgt), "joints_now")
    close(L.KeypointMSELoss(True)(pred, tgt, wgt), "kpmse_w")
    close(L.KeypointMSELoss(True)(pred, tgt, None), "kpmse_none")
    with pytest.raises(ValueError):
        L.FusedPoseLoss(True, "focal")
    with pytest.raises(ValueError):
        L.OffsetRegressionLoss("huber")


@pytest.mark.parametrize("name", NAMES)
def test_combined_loss_and_grads_against_golden(pkg, name):
    L = pkg.losses
    cfg, batch, ex, g = goldens_genb.load(name)
    config = types.SimpleNamespace(LOSS=types.SimpleNamespace(MORPH_LAMBDA=1.2, MORPH_WEIGHT=0.15, REG_WEIGHT=0.6))
    crit = L.build_loss(config)
    p = dev(ex["pred"]).requires_grad_(True)
    c = dev(ex["coords"]).requires_grad_(True)
    r = dev(ex["refined"]).requires_grad_(True)
    total, parts = crit({"heatmaps": p, "coords": c, "refined_coords": r},
                        {"heatmaps": dev(batch["target"]), "coords": dev(ex["target_coords"]), "weights": dev(batch["weight"])})
    total.backward()
    got = np.array([float(parts[k].detach()) for k in genb.COMBINED_KEYS])
    np.testing.assert_allclose(got, g["combined_f32"], rtol=LOSS_RTOL)
    maxnorm_close(p.grad.cpu().numpy(), g["combined_grad_pred"], what="d total / d pred")
    maxnorm_close(c.grad.cpu().numpy(), g["combined_grad_coords"], what="d total / d coords")
    maxnorm_close(r.grad.cpu().numpy(), g["combined_grad_refined"], what="d total / d refined")
    # the morphology term alone (its backward is the non-trivial one)
    p2 = dev(ex["pred"]).requires_grad_(True)
    L.MorphologyShapeLoss(1.2, 0.5)(p2, dev(batch["target"]), dev(batch["weight"])).backward()
    maxnorm_close(p2.grad.cpu().numpy(), g["morph_grad_pred"], what="d morph / d pred")


def test_combined_loss_arbitrary_upstream_and_sharding(pkg):
    """(1) upstream gradients that differ per term take the recompute path of the backward;
    (2) two half-batches with norm_batch = B give the full-batch gradients and losses that add up."""
    L = pkg.losses
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=33, B=8)
    ex = synth.make_genb_extras(cfg, batch, seed=33)
    config = types.SimpleNamespace(LOSS=types.SimpleNamespace(MORPH_LAMBDA=1.0, MORPH_WEIGHT=0.1, REG_WEIGHT=0.5))
    crit = L.CombinedLoss(config)

    def run_gpu(sl, norm_batch=0, mix=None):
        p = dev(ex["pred"][sl]).requires_grad_(True)
        c = dev(ex["coords"][sl]).requires_grad_(True)
        total, parts = crit({"heatmaps": p, "coords": c}, {"heatmaps": dev(batch["target"][sl]), "coords": dev(ex["target_coords"][sl]),
                                                           "weights": dev(batch["weight"][sl])}, norm_batch=norm_batch)
        obj = total if mix is None else mix(total, parts)
        obj.backward()
        return float(obj.detach()), p.grad.cpu().numpy(), c.grad.cpu().numpy(), {k: float(v.detach()) for k, v in parts.items()}

    def run_cpu(mix=None):
        p = T(ex["pred"]).clone().requires_grad_(True)
        c = T(ex["coords"]).clone().requires_grad_(True)
        total, parts = genb.combined_loss({"heatmaps": p, "coords": c}, {"heatmaps": T(batch["target"]), "coords": T(ex["target_coords"]),
                                                                         "weights": T(batch["weight"])}, 1.0, 0.1, 0.5)
        obj = total if mix is None else mix(total, parts)
        obj.backward()
        return float(obj.detach()), p.grad.numpy(), c.grad.numpy()

    mix = lambda total, parts: 3.0 * total + 2.0 * parts["heatmap"] - 0.5 * parts["morph"] + 0.25 * parts["regression"]
    full = slice(0, 8)
    got, want = run_gpu(full, mix=mix), run_cpu(mix=mix)
    np.testing.assert_allclose(got[0], want[0], rtol=LOSS_RTOL)
    maxnorm_close(got[1], want[1], what="mixed upstream, d/d pred")
    maxnorm_close(got[2], want[2], what="mixed upstream, d/d coords")

    whole = run_gpu(full)
    a, b = run_gpu(slice(0, 4), norm_batch=8), run_gpu(slice(4, 8), norm_batch=8)
    np.testing.assert_allclose(a[0] + b[0], whole[0], rtol=2e-6)
    for k in ("heatmap", "morph", "regression", "total"):
        np.testing.assert_allclose(a[3][k] + b[3][k], whole[3][k], rtol=2e-6)
    maxnorm_close(np.concatenate([a[1], b[1]]), whole[1], rel=1e-6, what="sharded d/d pred")
    maxnorm_close(np.concatenate([a[2], b[2]]), whole[2], rel=1e-6, what="sharded d/d coords")


def test_combined_loss_larger_batch_vs_oracle(pkg):
    L = pkg.losses
    cfg = synth.CONFIGS["hrformer_384x288"]
    batch = synth.make_batch(cfg, seed=41, B=24)
    ex = synth.make_genb_extras(cfg, batch, seed=41)
    config = types.SimpleNamespace(LOSS=types.SimpleNamespace(MORPH_LAMBDA=1.2, MORPH_WEIGHT=0.15, REG_WEIGHT=0.6))
    p = dev(ex["pred"]).requires_grad_(True)
    total, parts = L.CombinedLoss(config)({"heatmaps": p, "refined_coords": dev(ex["refined"])},
                                          {"heatmaps": dev(batch["target"]), "coords": dev(ex["target_coords"]), "weights": dev(batch["weight"])})
    total.backward()
    pc = T(ex["pred"]).clone().requires_grad_(True)
    tw, pw = genb.combined_loss({"heatmaps": pc, "refined_coords": T(ex["refined"])},
                                {"heatmaps": T(batch["target"]), "coords": T(ex["target_coords"]), "weights": T(batch["weight"])}, 1.2, 0.15, 0.6)
    tw.backward()
    assert set(parts) == set(pw) == {"heatmap", "morph", "refined", "total"}
    for k in parts:
        np.testing.assert_allclose(float(parts[k].detach()), float(pw[k].detach()), rtol=LOSS_RTOL, err_msg=k)
    maxnorm_close(p.grad.cpu().numpy(), pc.grad.numpy(), what="d total / d pred (B=24, 96x72)")


@pytest.mark.parametrize("name", NAMES + ["odd_56x40"])
def test_combined_loss_float16_predictions(pkg, name):
    """autocast: the head's heatmaps arrive in float16 (targets and coordinates stay float32).  The float16 entry points
    up-cast where they read and round the gradient once where they write, so against the float32 path on the up-cast
    maps the five losses are BIT-equal and the gradient is the float32 gradient rounded to half, bit for bit — with the
    loss scaler's factor as the assumed upstream (stored gradients used as they are), with an upstream that differs from
    the assumed one (recompute in the backward) and for the single-term modules.  Shapes: the three register-resident
    instantiations and one that takes the kernel's strided loop (56x40)."""
    L = pkg.losses
    if name == "odd_56x40":
        rng = np.random.default_rng(5)
        B, K, H, W = 3, 5, 56, 40
        pred32 = rng.normal(0.2, 0.5, (B, K, H, W)).astype(np.float32)
        target = np.clip(rng.normal(0.1, 0.3, (B, K, H, W)), 0, 1).astype(np.float32)
        weight = rng.integers(0, 3, (B, K)).astype(np.float32)
        coords = rng.uniform(0, 40, (B, K, 2)).astype(np.float32)
        tcoords = rng.uniform(0, 40, (B, K, 2)).astype(np.float32)
    else:
        cfg, batch, ex, g = goldens_genb.load(name)
        pred32, target, weight, coords, tcoords = ex["pred"], batch["target"], batch["weight"], ex["coords"], ex["target_coords"]
    config = types.SimpleNamespace(LOSS=types.SimpleNamespace(MORPH_LAMBDA=1.2, MORPH_WEIGHT=0.15, REG_WEIGHT=0.6))
    crit = L.build_loss(config)
    ph = dev(pred32).half()
    tg = {"heatmaps": dev(target), "coords": dev(tcoords), "weights": dev(weight)}

    def run(pred, scale_assumed, scale_actual):
        p = pred.clone().requires_grad_(True)
        c = dev(coords).requires_grad_(True)
        gs = None if scale_assumed is None else torch.tensor(float(scale_assumed)).cuda()
        total, parts = crit({"heatmaps": p, "coords": c}, tg, grad_scale=gs)
        (total * scale_actual).backward()
        return [float(parts[k].detach()) for k in ("heatmap", "morph", "regression", "total")], p.grad, c.grad

    for assumed, actual in ((None, 1.0), (65536.0, 65536.0), (None, 1024.0), (65536.0, 32768.0)):
        lh, gh, ch = run(ph, assumed, actual)
        lf, gf, cf = run(ph.float(), assumed, actual)
        assert gh.dtype == torch.float16 and gf.dtype == torch.float32
        assert lh == lf, f"{name} {assumed}/{actual}: losses differ {lh} vs {lf}"
        same(gh.cpu().numpy().view(np.uint16), gf.half().cpu().numpy().view(np.uint16), f"{name} {assumed}/{actual}: d/d pred (float16)")
        same(ch.cpu().numpy(), cf.cpu().numpy(), f"{name} {assumed}/{actual}: d/d coords")
        if actual == 1.0:                      # (the goldens' B = 1..3 make 65536 x gradient overflow half in both paths alike)
            assert torch.isfinite(gh.float()).all()
    # without the scale the small gradients flush to zero in half — the reason the scale must be met BEFORE rounding
    _, g1, _ = run(ph, None, 1.0)
    _, g2, _ = run(ph, 65536.0, 65536.0)
    assert (g2 != 0).float().mean() >= (g1 != 0).float().mean()
    # single-term modules, the criterion variants and a forward under no_grad
    for mod, args in ((L.FusedPoseLoss(True, "smoothl1"), (tg["heatmaps"], tg["weights"])), (L.MorphologyShapeLoss(1.2, 0.5), (tg["heatmaps"], tg["weights"])),
                      (L.JointsMSELoss(True), (tg["heatmaps"], tg["weights"])), (L.KeypointMSELoss(True), (tg["heatmaps"], tg["weights"]))):
        a = ph.clone().requires_grad_(True)
        b = ph.float().requires_grad_(True)
        la, lb = mod(a, *args), mod(b, *args)
        assert float(la) == float(lb), type(mod).__name__
        (la * 4096.0).backward(); (lb * 4096.0).backward()
        same(a.grad.cpu().numpy().view(np.uint16), b.grad.half().cpu().numpy().view(np.uint16), type(mod).__name__)
        with torch.no_grad():
            assert float(mod(ph, *args)) == float(lb)
    # a slice of whole images of a float16 batch starts on an 8-byte boundary, not necessarily a 16-byte one
    if name == "odd_56x40":
        small = torch.rand(3, 1, 3, 4, device="cuda").half()
        tsm = torch.rand(3, 1, 3, 4, device="cuda")
        view = small[1:]
        assert view.data_ptr() % 16 == 8 and view.is_contiguous()
        a = view.clone().requires_grad_(True)          # (a clone is 16-byte aligned again: the reference value)
        la = L.KeypointMSELoss(False)(a, tsm[1:], None)
        la.backward()
        with torch.no_grad():
            assert float(L.KeypointMSELoss(False)(view, tsm[1:], None)) == float(la)
        b = view.detach().requires_grad_(True)
        lb = L.KeypointMSELoss(False)(b, tsm[1:], None)
        lb.backward()
        assert float(lb) == float(la) and torch.equal(a.grad, b.grad)
    # torch.autocast hands the same dtype over: the module is called inside the region as train.py does
    with torch.autocast("cuda", dtype=torch.float16):
        a = ph.clone().requires_grad_(True)
        total, _ = crit({"heatmaps": a, "coords": dev(coords)}, tg)
    total.backward()
    lf, gf, _ = run(ph.float(), None, 1.0)
    assert float(total) == lf[3]
    same(a.grad.cpu().numpy().view(np.uint16), gf.half().cpu().numpy().view(np.uint16), "under torch.autocast")


# ------------------------------------------------------------------ encoders
@pytest.mark.parametrize("name", NAMES)
def test_encoders_against_golden(pkg, name):
    gh = pkg.generate_heatmap
    cfg, batch, ex, g = goldens_genb.load(name)
    W, H = cfg.heatmap_size
    t, w = gh.generate_heatmaps_clipped(dev(batch["kps"]), dev(batch["vis"]), (H, W), cfg.input_size, cfg.sigma)
    same(w.cpu().numpy(), g["clip_weight"], "clipped weights")
    # patch values: CUDA's expf and numpy's float32 exp may differ in the last bit (same tolerance as the
    # Gen-A encoder, 1e-6 relative); the pasted support must be identical
    same(t.cpu().numpy() != 0, g["clip_target"] != 0, "clipped tiles, support")
    np.testing.assert_allclose(t.cpu().numpy(), g["clip_target"], rtol=1e-6, atol=0)
    ek, ev = synth.edge_keypoints(cfg)
    t, w = gh.generate_heatmaps_clipped(dev(ek), dev(ev), (H, W), cfg.input_size, cfg.sigma)
    same(w.cpu().numpy(), g["clip_edge_weight"], "clipped weights, edge set")
    same(t.cpu().numpy() != 0, g["clip_edge_target"] != 0, "clipped tiles, edge set, support")
    np.testing.assert_allclose(t.cpu().numpy(), g["clip_edge_target"], rtol=1e-6, atol=0)
    gen = gh.GenerateTarget({"input_size": (cfg.input_size[1], cfg.input_size[0]), "heatmap_size": (H, W), "sigma": cfg.sigma})
    r = gen({"keypoints": dev(batch["kps"]), "keypoints_visible": dev(batch["vis"])})
    assert np.array_equal(r["keypoint_weights"].cpu().numpy(), g["dense_weight"])
    # float32 exp: numpy's and CUDA's expf may differ in the last bit; sub-normal results get an absolute floor
    np.testing.assert_allclose(r["heatmaps"][:1].cpu().numpy(), g["dense0"], rtol=1e-6, atol=1e-37)
    r = gen({"keypoints": dev(ek[:1]), "keypoints_visible": dev(ev[:1])})
    np.testing.assert_allclose(r["heatmaps"].cpu().numpy(), g["dense_edge0"], rtol=1e-6, atol=1e-37)
    assert np.array_equal(r["keypoint_weights"].cpu().numpy(), g["dense_edge0_weight"])


def test_generate_target_example_of_the_reference(pkg):
    """data/test_transforms.py:342-379: peak 1.0 at (row, col) (32,24), (30,25), (35,20) of a 48x64 map."""
    g = goldens_genb.load_test_transforms()
    gen = pkg.generate_heatmap.GenerateTarget({"input_size": (192, 256), "heatmap_size": (48, 64), "sigma": 2.0})
    r = gen({"keypoints": dev(g["keypoints"]), "keypoints_visible": torch.ones(3).cuda()})
    h = r["heatmaps"].cpu().numpy()
    assert h.shape == (3, 48, 64)
    np.testing.assert_allclose(h, g["heatmaps"], rtol=1e-6, atol=1e-37)
    peaks = [(float(t.max()), *np.unravel_index(t.argmax(), t.shape)) for t in h]
    assert peaks == [(1.0, 32, 24), (1.0, 30, 25), (1.0, 35, 20)]


def test_encoders_full_size_properties(pkg):
    """BASELINE configs[1] size (B=1024): properties that need no oracle run — every active clipped tile
    holds exactly one 1.0 at (ul_c + centre); the dense tile's maximum sits at round(centre) and is symmetric
    to the patch encoder for integer centres."""
    from infantposeestimation_gaussianbias_b200 import _native as N, ops
    cfg = synth.CONFIGS["w32_256x192"]
    rng = np.random.default_rng(7)
    kps, vis = synth.make_keypoints(cfg, rng, 1024)
    W, H = cfg.heatmap_size
    t, w = ops.encode_mode(dev(kps), dev(vis), H, W, 192.0, 256.0, 2.0, N.ENCODE_PATCH_CLIPPED)
    mu = kps * np.float32(0.25)
    inside = (vis > 0) & (mu[..., 0] >= 0) & (mu[..., 1] >= 0) & (mu[..., 0] < W) & (mu[..., 1] < H)
    assert np.array_equal(w.cpu().numpy()[..., 0], inside.astype(np.float32))
    tmax = t.amax(dim=(2, 3)).cpu().numpy()
    assert np.all(tmax[~inside] == 0)
    # the clamped origin can push the patch centre out of the pasted block only when br - ul_c <= centre
    assert np.all((tmax[inside] <= 1.0) & (tmax[inside] > 0))
    d, wd = ops.encode_mode(dev(np.round(kps / 4) * 4), dev(vis), H, W, 192.0, 256.0, 2.0, N.ENCODE_DENSE)
    assert np.array_equal(wd.cpu().numpy()[..., 0], ((vis > 0) & (np.round(kps / 4)[..., 0] >= 0) & (np.round(kps / 4)[..., 1] >= 0)
                                                    & (np.round(kps / 4)[..., 0] < W) & (np.round(kps / 4)[..., 1] < H)).astype(np.float32))
    dm = d.amax(dim=(2, 3)).cpu().numpy()
    assert np.all(dm[wd.cpu().numpy()[..., 0] > 0] == 1.0)


def test_errors_are_loud(pkg):
    from infantposeestimation_gaussianbias_b200 import GbcodecError, _native as N, ops
    with pytest.raises(RuntimeError):
        ops.encode_mode(torch.zeros(1, 1, 2), torch.ones(1, 1), 8, 8, 32.0, 32.0, 2.0, 1)       # CPU tensors
    with pytest.raises(GbcodecError):
        ops.encode_mode(torch.zeros(1, 1, 2).cuda(), torch.ones(1, 1).cuda(), 8, 8, 32.0, 32.0, 2.0, 9)   # bad mode
    with pytest.raises(GbcodecError):
        ops.postprocess(torch.zeros(1, 1, 8, 8).cuda(), None, None, None, N.ARGMAX_TAYLOR, False, 256.0, 5, True, 0.3, True, 256.0, 256.0)   # transform without centre


# ------------------------------------------------------------------ plain heatmap head, one pass
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("on_the_fly", [False, True])
def test_heatmap_head_step(pkg, name, on_the_fly):
    """KeypointMSELoss fwd + bwd with (optionally on-the-fly) targets + decode_heatmaps in one kernel, against the
    golden KeypointMSELoss value of the reference and the oracle's gradient / arg-max."""
    from oracle import heatmap_codec as oc
    cfg, batch, ex, g = goldens_genb.load(name)
    step = pkg.pose_estimator.HeatmapHeadStep(input_size=cfg.input_size, sigma=cfg.sigma)
    p = dev(ex["pred"]).requires_grad_(True)
    if on_the_fly:
        loss, kp, mv = step(p, keypoints=dev(batch["kps"]), keypoints_visible=dev(batch["vis"]))
    else:
        loss, kp, mv = step(p, dev(batch["target"]), dev(batch["weight"]))
    (loss * 3.0).backward()
    np.testing.assert_allclose(float(loss.detach()), float(g["kpmse_w"]), rtol=LOSS_RTOL)
    pc = T(ex["pred"]).clone().requires_grad_(True)
    (genb.keypoint_mse_loss(pc, T(batch["target"]), T(batch["weight"])) * 3.0).backward()
    maxnorm_close(p.grad.cpu().numpy(), pc.grad.numpy(), what="d KeypointMSELoss / d pred")
    wkp, wmv, _ = oc.decode_heatmaps(T(ex["pred"]), shift=True)
    assert np.array_equal(kp.cpu().numpy(), wkp.numpy()) and np.array_equal(mv.cpu().numpy(), wmv.numpy())
    # no weights: plain mean squared error
    l2 = pkg.pose_estimator.HeatmapHeadStep(input_size=cfg.input_size, sigma=cfg.sigma)(dev(ex["pred"]), dev(batch["target"]), None, decode=False)
    np.testing.assert_allclose(float(l2), float(g["kpmse_none"]), rtol=LOSS_RTOL)


def test_heatmap_head_step_sharded_and_ties(pkg):
    from oracle import heatmap_codec as oc
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=61, B=8)
    step = pkg.pose_estimator.HeatmapHeadStep(input_size=cfg.input_size, sigma=cfg.sigma)
    hm = batch["heatmaps"]                      # includes constant tiles and duplicated maxima (first-max rule)
    def run(sl, nb=0):
        p = dev(hm[sl]).requires_grad_(True)
        loss, kp, mv = step(p, keypoints=dev(batch["kps"][sl]), keypoints_visible=dev(batch["vis"][sl]), norm_batch=nb)
        loss.backward()
        return float(loss.detach()), p.grad.cpu().numpy(), kp.cpu().numpy(), mv.cpu().numpy()
    whole = run(slice(0, 8))
    a, b = run(slice(0, 3), 8), run(slice(3, 8), 8)
    np.testing.assert_allclose(a[0] + b[0], whole[0], rtol=2e-6)
    assert np.array_equal(np.concatenate([a[1], b[1]]), whole[1])
    wkp, wmv, _ = oc.decode_heatmaps(T(hm), shift=True)
    assert np.array_equal(whole[2], wkp.numpy()) and np.array_equal(whole[3], wmv.numpy())
    want = genb.keypoint_mse_loss(T(hm), T(batch["target"]), T(batch["weight"]))
    np.testing.assert_allclose(whole[0], float(want), rtol=LOSS_RTOL)
