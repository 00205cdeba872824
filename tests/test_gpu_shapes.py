"""GPU: tile shapes outside the three instantiated ones (64x48, 96x72, 128x128) take the shape-agnostic
kernels (loss_kernel, decode_kernel<0>, argmax_kernel<0>, genb_tile_kernel<0>, postprocess_kernel<0>) or other
register-resident instantiations.  Same oracle, same tolerances.  Shapes: the reference's other default
(64x64, data/pose_transforms.py:391), a small map, two whose float4 count has no whole-warp divisor <= 8 passes
(56x40, 80x60), the largest tile the ABI accepts (128x256 = GBCODEC_MAX_TILE), K from 1 to 17, sigma 1 to 3."""
import types

import numpy as np
import pytest
import torch

from oracle import genb, heatmap_codec as oc
from tests import synth

pytestmark = pytest.mark.gpu
ENC_RTOL, LOSS_RTOL, COORD_ATOL = 1e-6, 1e-5, 1e-4

SHAPES = {
    "64x64": synth.CodecConfig("64x64", 3, 17, (64, 64), (256, 256), 2.0),
    "32x24": synth.CodecConfig("32x24", 4, 17, (24, 32), (96, 128), 1.0),
    "56x40": synth.CodecConfig("56x40", 3, 5, (40, 56), (160, 224), 2.0),
    "80x60": synth.CodecConfig("80x60", 2, 17, (60, 80), (240, 320), 3.0),
    "128x256": synth.CodecConfig("128x256", 1, 3, (256, 128), (512, 256), 2.0),       # GBCODEC_MAX_TILE
    "128x192": synth.CodecConfig("128x192", 1, 3, (192, 128), (384, 256), 2.0),       # largest family the six-term loss takes
    "16x16_K1": synth.CodecConfig("16x16_K1", 5, 1, (16, 16), (64, 64), 1.0),
}


@pytest.fixture(scope="module")
def pkg():
    import infantposeestimation_gaussianbias_b200 as p
    p.load()
    return p


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def maxnorm_close(got, want, rel=LOSS_RTOL, what=""):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale, err = np.abs(want).max(), np.abs(got - want).max()
    assert err <= rel * scale + 1e-30, f"{what}: max err {err:.3e} vs scale {scale:.3e} ({err / max(scale, 1e-300):.2e})"


@pytest.mark.parametrize("name", list(SHAPES))
def test_encode_and_decode(pkg, name):
    cfg = SHAPES[name]
    ops = pkg.ops
    batch = synth.make_batch(cfg, seed=17)
    t, w = ops.encode(dev(batch["kps"]), dev(batch["vis"]), cfg.H, cfg.W, float(cfg.input_size[0]), float(cfg.input_size[1]), cfg.sigma)
    assert np.array_equal(w.cpu().numpy(), batch["weight"]) and np.array_equal(t.cpu().numpy() != 0, batch["target"] != 0)
    np.testing.assert_allclose(t.cpu().numpy(), batch["target"], rtol=ENC_RTOL, atol=0)
    ek, ev = synth.edge_keypoints(cfg)
    t, w = ops.encode(dev(ek), dev(ev), cfg.H, cfg.W, float(cfg.input_size[0]), float(cfg.input_size[1]), cfg.sigma)
    ot, ow = oc.encode_targets(ek, ev, cfg.heatmap_size, cfg.input_size, cfg.sigma)
    assert np.array_equal(w.cpu().numpy(), ow) and np.array_equal(t.cpu().numpy() != 0, ot != 0)
    # arg-max family: bit-exact
    hm = T(batch["heatmaps"])
    for mode, shift in ((0, False), (1, True)):
        c, v, _ = ops.decode_argmax(hm.cuda(), mode)
        wc, wv, _ = oc.decode_heatmaps(hm, shift=shift)
        assert np.array_equal(c.cpu().numpy(), wc.numpy()) and np.array_equal(v.cpu().numpy(), wv.numpy())
    c, _, _ = ops.decode_argmax(hm.cuda(), 2)
    assert np.array_equal(c.cpu().numpy(), genb.get_max_preds_with_subpixel(hm)[0].numpy())
    # fusion decode with and without the flip average
    a, f = torch.tensor([0.3]).cuda(), torch.tensor([0.7]).cuda()
    c, s, centre = ops.decode(hm.cuda(), None, None, dev(batch["offsets"]), a, f, 2, 3)
    want_c, want_s = oc.fusion_decode(hm, T(batch["offsets"]), 0.3, 0.7, centre=T(centre.cpu().numpy()))
    assert np.array_equal(s.cpu().numpy(), want_s.numpy())
    assert np.abs(c.cpu().numpy() - want_c.numpy()).max() <= COORD_ATOL
    glob = oc.soft_argmax(hm)[0].numpy()
    ok = (np.abs(glob - np.floor(glob) - 0.5) > 1e-3).all(-1)
    assert np.array_equal(centre.cpu().numpy()[ok], oc.window_centres(T(glob), cfg.H, cfg.W).numpy()[ok])
    if cfg.K == 17:
        perm = synth.flip_perm(cfg.K)
        cf, sf, cen = ops.decode(hm.cuda(), dev(batch["heatmaps_flip"]), dev(perm), dev(batch["offsets"]), a, f, 2, 3)
        avg = oc.flip_average(hm, T(batch["heatmaps_flip"]))
        want_c, want_s = oc.fusion_decode(avg, T(batch["offsets"]), 0.3, 0.7, centre=T(cen.cpu().numpy()))
        assert np.array_equal(sf.cpu().numpy(), want_s.numpy())
        assert np.abs(cf.cpu().numpy() - want_c.numpy()).max() <= COORD_ATOL


@pytest.mark.parametrize("name", list(SHAPES))
@pytest.mark.parametrize("on_the_fly", [False, True])
def test_fusion_loss(pkg, name, on_the_fly):
    cfg = SHAPES[name]
    ops = pkg.ops
    batch = synth.make_batch(cfg, seed=18)
    pairs = [v for p in oc.skeleton_for(cfg.K) for v in p]
    if name == "128x256":
        # the loss keeps two tiles in shared memory: 2 x 128 KB does not fit, and it says so
        from infantposeestimation_gaussianbias_b200 import GbcodecError
        with pytest.raises(GbcodecError, match="shared memory"):
            ops.fusion_loss(dev(batch["heatmaps"]), dev(batch["offsets"]), dev(batch["variances"]), dev(batch["target"]), dev(batch["weight"]),
                            dev(batch["kps"]), None, None, float(cfg.input_size[0]), float(cfg.input_size[1]), list(oc.DEFAULT_LAMBDAS),
                            cfg.sigma, cfg.sigma, True, pairs, True, False, None, None, 2, 0)
        return
    res = ops.fusion_loss(dev(batch["heatmaps"]), dev(batch["offsets"]), dev(batch["variances"]), None if on_the_fly else dev(batch["target"]),
                          dev(batch["vis"][..., None] if on_the_fly else batch["weight"]), dev(batch["kps"]), None, None,
                          float(cfg.input_size[0]), float(cfg.input_size[1]), list(oc.DEFAULT_LAMBDAS), cfg.sigma, cfg.sigma, True, pairs,
                          True, False, None, None, 2, 0)
    want_l, want_g = oc.fusion_loss_and_grads(T(batch["heatmaps"]), T(batch["offsets"]), T(batch["variances"]), T(batch["target"]),
                                              T(batch["weight"]), T(batch["kps"]), input_size=cfg.input_size, target_sigma=cfg.sigma)
    for i, k in enumerate(oc.LOSS_KEYS):
        np.testing.assert_allclose(float(res[0][i]), float(want_l[k]), rtol=LOSS_RTOL, atol=1e-9, err_msg=k)
    maxnorm_close(res[1].cpu().numpy(), want_g["heatmaps"].numpy(), what="grad heatmaps")
    maxnorm_close(res[3].cpu().numpy(), want_g["variances"].numpy(), what="grad variances")
    # the offset gradient inherits the soft-argmax's fp32 noise: compare against fp64 truth with the reference's own error bar
    _, g64 = oc.fusion_loss_and_grads(T(batch["heatmaps"]).double(), T(batch["offsets"]).double(), T(batch["variances"]).double(),
                                      T(batch["target"]).double(), T(batch["weight"]).double(), T(batch["kps"]).double(),
                                      input_size=cfg.input_size, target_sigma=cfg.sigma)
    truth, ref32, got = g64["offsets"].numpy(), want_g["offsets"].numpy().astype(np.float64), res[2].cpu().numpy().astype(np.float64)
    scale = np.abs(truth).max()
    assert np.abs(got - truth).max() <= max(LOSS_RTOL * scale, 3 * np.abs(ref32 - truth).max())


@pytest.mark.parametrize("name", list(SHAPES))
def test_genb_family(pkg, name):
    cfg = SHAPES[name]
    batch = synth.make_batch(cfg, seed=19)
    ex = synth.make_genb_extras(cfg, batch, seed=19)
    config = types.SimpleNamespace(LOSS=types.SimpleNamespace(MORPH_LAMBDA=1.2, MORPH_WEIGHT=0.15, REG_WEIGHT=0.6))
    p = dev(ex["pred"]).requires_grad_(True)
    total, parts = pkg.losses.CombinedLoss(config)({"heatmaps": p, "coords": dev(ex["coords"])},
                                                   {"heatmaps": dev(batch["target"]), "coords": dev(ex["target_coords"]), "weights": dev(batch["weight"])})
    total.backward()
    pc = T(ex["pred"]).clone().requires_grad_(True)
    tw, pw = genb.combined_loss({"heatmaps": pc, "coords": T(ex["coords"])},
                                {"heatmaps": T(batch["target"]), "coords": T(ex["target_coords"]), "weights": T(batch["weight"])}, 1.2, 0.15, 0.6)
    tw.backward()
    for k in pw:
        np.testing.assert_allclose(float(parts[k].detach()), float(pw[k].detach()), rtol=LOSS_RTOL, err_msg=k)
    maxnorm_close(p.grad.cpu().numpy(), pc.grad.numpy(), what="d total / d pred")
    # the whole postprocess pipeline
    reg = np.abs(ex["reg_px"])
    cfg_pp = types.SimpleNamespace(TEST=types.SimpleNamespace())
    r = pkg.postprocess.postprocess_predictions({"heatmaps": dev(ex["pred"]), "coords": dev(reg)}, {"center": dev(ex["center"]), "scale": dev(ex["scale"])}, cfg_pp)
    want = genb.postprocess_predictions(T(ex["pred"]), T(reg), T(ex["center"]), T(ex["scale"]))
    assert np.array_equal(r["mask"].cpu().numpy(), want["mask"].numpy())
    np.testing.assert_allclose(r["preds"].cpu().numpy(), want["preds"].numpy(), rtol=2.5e-7, atol=COORD_ATOL)
    # the two second-generation encoders
    gh = pkg.generate_heatmap
    t, w = gh.generate_heatmaps_clipped(dev(batch["kps"]), dev(batch["vis"]), (cfg.H, cfg.W), cfg.input_size, cfg.sigma)
    ot, ow = genb.encode_patch_clipped(batch["kps"], batch["vis"], (cfg.H, cfg.W), cfg.input_size, cfg.sigma)
    assert np.array_equal(w.cpu().numpy(), ow) and np.array_equal(t.cpu().numpy() != 0, ot != 0)
    np.testing.assert_allclose(t.cpu().numpy(), ot, rtol=ENC_RTOL, atol=0)
    gen = gh.GenerateTarget({"input_size": (cfg.input_size[1], cfg.input_size[0]), "heatmap_size": (cfg.H, cfg.W), "sigma": cfg.sigma})
    rr = gen({"keypoints": dev(batch["kps"]), "keypoints_visible": dev(batch["vis"])})
    oh, ow = genb.encode_dense(batch["kps"], batch["vis"], (cfg.H, cfg.W), (cfg.input_size[1], cfg.input_size[0]), cfg.sigma)
    assert np.array_equal(rr["keypoint_weights"].cpu().numpy(), ow)
    np.testing.assert_allclose(rr["heatmaps"].cpu().numpy(), oh, rtol=ENC_RTOL, atol=1e-37)


def test_limits_are_enforced(pkg):
    from infantposeestimation_gaussianbias_b200 import GbcodecError
    ops = pkg.ops
    with pytest.raises(GbcodecError):
        ops.decode_argmax(torch.zeros(1, 1, 256, 132).cuda(), 0)          # tile above GBCODEC_MAX_TILE
    with pytest.raises(GbcodecError):
        ops.decode_argmax(torch.zeros(1, 65, 8, 8).cuda(), 0)             # K above GBCODEC_MAX_K
    with pytest.raises(GbcodecError):
        ops.encode(torch.zeros(1, 1, 2).cuda(), torch.ones(1, 1).cuda(), 8, 8, 32.0, 32.0, 0.0)    # sigma


def test_64x64_takes_the_tile_kernels_in_every_mode(pkg):
    """64x64 (the reference's other default map, data/pose_transforms.py:391) has its own instantiations: float32, float16
    maps and per-tile variance means must agree with each other there too."""
    cfg = SHAPES["64x64"]
    ops = pkg.ops
    batch = synth.make_batch(cfg, seed=29)
    pairs = [v for p in oc.skeleton_for(cfg.K) for v in p]
    a, f = torch.tensor(0.5).cuda(), torch.tensor(0.62).cuda()
    common = (float(cfg.input_size[0]), float(cfg.input_size[1]), list(oc.DEFAULT_LAMBDAS), cfg.sigma, cfg.sigma, True, pairs)
    half = {k: dev(batch[k]).half() for k in ("heatmaps", "offsets", "variances")}
    up = {k: v.float() for k, v in half.items()}
    full = ops.fusion_loss(up["heatmaps"], up["offsets"], up["variances"], None, dev(batch["vis"]), dev(batch["kps"]), None, None, *common, True, True, a, f, 2, 3)
    h16 = ops.fusion_loss_f16(half["heatmaps"], half["offsets"], half["variances"], None, dev(batch["vis"]), dev(batch["kps"]), None, None, *common,
                              True, True, a, f, 2, 3)
    assert torch.equal(h16[0], full[0]) and torch.equal(h16[1], full[4]) and torch.equal(h16[2], full[5])
    assert (h16[3] == full[1].half()).float().mean().item() > 0.9999
    vm = up["variances"].double().mean(dim=(2, 3)).float()
    res = ops.fusion_step_vmean(up["heatmaps"], up["offsets"], vm, None, dev(batch["vis"]), dev(batch["kps"]), None, None, *common, True, True, a, f, 2, 3)
    np.testing.assert_allclose(res[0].cpu().numpy(), full[0].cpu().numpy(), rtol=2e-6, atol=1e-9)
    assert torch.equal(res[1], full[1]) and torch.equal(res[2], full[2])


@pytest.mark.parametrize("name", ["hrformer_384x288", "preemie_256"])
def test_float16_maps_on_the_ring_instantiations(pkg, name):
    """96x72 and 128x128 have no whole-tile partner slots: partner rows (and at 128x128 the variance rows) stream through a
    two-row ring.  With float16 maps the ring carries raw halves; losses, decode and gradients must still be those of the
    float32 path on the up-cast maps (gradients rounded once to half, up to the tie pixels that are stored twice)."""
    cfg = synth.CONFIGS[name]
    ops = pkg.ops
    batch = synth.make_batch(cfg, seed=31, B=5)
    pairs = [v for p in oc.skeleton_for(cfg.K) for v in p]
    a, f = torch.tensor(0.5).cuda(), torch.tensor(0.62).cuda()
    common = (float(cfg.input_size[0]), float(cfg.input_size[1]), list(oc.DEFAULT_LAMBDAS), cfg.sigma, cfg.sigma, True, pairs)
    half = {k: dev(batch[k]).half() for k in ("heatmaps", "offsets", "variances")}
    up = {k: v.float() for k, v in half.items()}
    full = ops.fusion_loss(up["heatmaps"], up["offsets"], up["variances"], None, dev(batch["vis"]), dev(batch["kps"]), None, None, *common, True, True, a, f, 2, 3)
    h16 = ops.fusion_loss_f16(half["heatmaps"], half["offsets"], half["variances"], None, dev(batch["vis"]), dev(batch["kps"]), None, None, *common,
                              True, True, a, f, 2, 3)
    assert torch.equal(h16[0], full[0]) and torch.equal(h16[1], full[4]) and torch.equal(h16[2], full[5])
    assert (h16[3] == full[1].half()).float().mean().item() > 0.9999
    assert torch.equal(h16[4], full[2].half()) and torch.equal(h16[5], full[3].half())
