"""CPU: pin the oracle restatement to the outputs of the reference itself
(tests/golden/*.npz, see tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import heatmap_codec as oc
from tests import goldens, synth

NAMES = list(synth.CONFIGS)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("name", NAMES)
def test_encode_bit_equal(name):
    cfg, batch, g = goldens.load(name)
    target, weight = oc.encode_targets(batch["kps"], batch["vis"], cfg.heatmap_size, cfg.input_size, cfg.sigma)
    assert np.array_equal(target, g["target"])
    assert np.array_equal(weight, g["enc_weight"])
    target, weight = oc.encode_targets(g["edge_kps"], g["edge_vis"], cfg.heatmap_size, cfg.input_size, cfg.sigma)
    assert np.array_equal(target, g["edge_target"])
    assert np.array_equal(weight, g["edge_weight"])


def test_encode_known_quirks():
    # SURVEY Q2-Q4, each checked against the reference at survey time.
    hs, ins = (48, 64), (192, 256)
    kp = np.array([[[-7.5 * 4, 20 * 4], [1.5 * 4, 4.2 * 4], [100 * 4, 10 * 4], [-0.3 * 4, 0.0]]], np.float32)
    vis = np.array([[2, 1, 2, 0]], np.float32)
    target, weight = oc.encode_targets(kp, vis, hs, ins, 2.0)
    # br == 0: nothing pasted, weight kept
    assert target[0, 0].max() == 0 and weight[0, 0, 0] == 2
    # peak sits on trunc(mu-6)+6: mu=1.5 -> int(-4.5)=-4 -> 2 ; mu=4.2 -> int(-1.8)=-1 -> 5
    assert np.unravel_index(target[0, 1].argmax(), (64, 48)) == (5, 2) and target[0, 1].max() == 1.0
    # wholly off-map: weight zeroed
    assert weight[0, 2, 0] == 0 and target[0, 2].max() == 0
    # invisible: untouched
    assert weight[0, 3, 0] == 0 and target[0, 3].max() == 0
    # sigma 1.5: 10 taps, centre 5
    target, _ = oc.encode_targets(np.array([[[20.0 * 2, 30.0 * 2]]], np.float32), np.array([[1.0]], np.float32),
                                  (128, 128), (256, 256), 1.5)
    ys, xs = np.nonzero(target[0, 0])
    assert (xs.min(), xs.max(), ys.min(), ys.max()) == (15, 24, 25, 34)
    assert np.unravel_index(target[0, 0].argmax(), (128, 128)) == (30, 20)


@pytest.mark.parametrize("name", NAMES)
def test_loss_and_grads(name):
    cfg, batch, g = goldens.load(name)
    losses, grads = oc.fusion_loss_and_grads(
        t(batch["heatmaps"]), t(batch["offsets"]), t(batch["variances"]), t(batch["target"]),
        t(batch["weight"]), t(batch["kps"]), input_size=cfg.input_size, target_sigma=cfg.sigma)
    got = np.array([float(losses[k]) for k in oc.LOSS_KEYS])
    np.testing.assert_allclose(got, g["loss_f32"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(got, g["loss_f64"], rtol=1e-5, atol=1e-9)
    gh = grads["heatmaps"].numpy()
    scale = np.abs(g["grad_hm"]).max()
    assert np.abs(gh - g["grad_hm"]).max() <= 2e-6 * scale
    go = grads["offsets"].numpy()
    assert np.array_equal(go != 0, g["grad_off"] != 0)
    np.testing.assert_allclose(go, g["grad_off"], rtol=1e-5, atol=1e-12)
    gv = grads["variances"].numpy()
    np.testing.assert_allclose(gv, np.broadcast_to(g["grad_var_tile"][:, :, None, None], gv.shape), rtol=1e-5, atol=1e-14)


@pytest.mark.parametrize("name", NAMES)
def test_loss_f64_matches_reference_f64(name):
    cfg, batch, g = goldens.load(name)
    d = lambda k: t(batch[k]).double()
    losses = oc.fusion_loss(d("heatmaps"), d("offsets"), d("variances"), d("target"), d("weight"), d("kps"),
                            input_size=cfg.input_size, target_sigma=cfg.sigma)
    got = np.array([float(losses[k]) for k in oc.LOSS_KEYS])
    np.testing.assert_allclose(got, g["loss_f64"], rtol=1e-12)
    _, grads = oc.fusion_loss_and_grads(d("heatmaps"), d("offsets"), d("variances"), d("target"), d("weight"), d("kps"),
                                        input_size=cfg.input_size, target_sigma=cfg.sigma)
    np.testing.assert_allclose(grads["heatmaps"].numpy().reshape(-1)[::97], g["grad_hm_f64_sub"], rtol=1e-9, atol=1e-18)
    np.testing.assert_allclose(grads["offsets"].numpy(), g["grad_off_f64"], rtol=1e-9, atol=1e-18)
    np.testing.assert_allclose(grads["variances"].numpy()[:, :, 0, 0], g["grad_var_f64_tile"], rtol=1e-9, atol=1e-18)


@pytest.mark.parametrize("name", NAMES)
def test_decode(name):
    cfg, batch, g = goldens.load(name)
    hm, off = t(batch["heatmaps"]), t(batch["offsets"])
    sa, sc = oc.soft_argmax(hm)
    np.testing.assert_allclose(sa.numpy(), g["dec_softargmax"], rtol=0, atol=2e-5)
    assert np.array_equal(sc.numpy(), g["dec_scores"])
    for loop in (False, True):
        c, s = oc.fusion_decode(hm, off, float(g["alpha_param"]), float(g["fusion_weight"]), loop=loop)
        assert np.abs(c.numpy() - g["dec_coords"]).max() <= 1e-4
        assert np.array_equal(s.numpy(), g["dec_scores"])
    c, _ = oc.fusion_decode(hm, off, float(g["alpha_param"]), float(g["fusion_weight"]), apply_offset=False)
    assert np.abs(c.numpy() - g["dec_coords_nooff"]).max() <= 1e-4
    pairs = [p for p in oc.COCO_FLIP_PAIRS if p[0] < cfg.K and p[1] < cfg.K]
    c, s = oc.fusion_decode(hm, off, float(g["alpha_param"]), float(g["fusion_weight"]),
                            heatmaps_of_flipped_input=t(batch["heatmaps_flip"]), flip_pairs=pairs)
    assert np.abs(c.numpy() - g["flip_coords"]).max() <= 1e-4
    assert np.array_equal(s.numpy(), g["flip_scores"])


@pytest.mark.parametrize("name", NAMES)
def test_decode_heatmaps_bit_exact(name):
    cfg, batch, g = goldens.load(name)
    c, v, idx = oc.decode_heatmaps(t(batch["heatmaps"]), shift=True)
    assert np.array_equal(idx.numpy(), g["argmax_idx"])
    assert np.array_equal(v.numpy(), g["argmax_vals"])
    assert np.array_equal(c.numpy(), g["argmax_coords"])


def test_first_max_tie_break():
    h = torch.zeros(1, 2, 4, 8)
    h[0, 1, 2, 3] = 1.0
    h[0, 1, 3, 1] = 1.0
    c, v, idx = oc.decode_heatmaps(h, shift=True)
    assert idx.tolist() == [[0, 19]]
    assert c[0, 0].tolist() == [0.0, 0.0]


def test_sharded_denominators_reproduce_global_batch():
    cfg, batch, g = goldens.load("w32_256x192")
    full = oc.fusion_loss(t(batch["heatmaps"]), t(batch["offsets"]), t(batch["variances"]), t(batch["target"]),
                          t(batch["weight"]), t(batch["kps"]), input_size=cfg.input_size)
    den = oc.loss_denominators(t(batch["weight"]), cfg.K)
    acc = None
    for sl in (slice(0, 1), slice(1, 3)):
        part = oc.fusion_loss(*(t(batch[k][sl]) for k in ("heatmaps", "offsets", "variances", "target", "weight", "kps")),
                              input_size=cfg.input_size, denominators=den)
        acc = part if acc is None else {k: acc[k] + part[k] for k in acc}
    for k in oc.LOSS_KEYS:
        assert abs(float(acc[k]) - float(full[k])) <= 2e-6 * abs(float(full[k])) + 1e-9


@pytest.mark.parametrize("name", NAMES)
def test_oracle_against_large_goldens(name):
    """The second golden set (BASELINE configs[0] size: B = 32 at 64x48, B = 8 for the larger shapes), produced by the
    reference itself: the oracle must reproduce it too — encode bit for bit, losses / gradients / decode tightly."""
    from tests.golden.make_golden import LARGE_STRIDE, tile_moments
    cfg, batch, g = goldens.load_large(name)
    target, weight = oc.encode_targets(batch["kps"], batch["vis"], cfg.heatmap_size, cfg.input_size, cfg.sigma)
    assert np.array_equal(target, g["target"]) and np.array_equal(weight, g["enc_weight"])
    losses, grads = oc.fusion_loss_and_grads(t(batch["heatmaps"]), t(batch["offsets"]), t(batch["variances"]), t(target), t(weight),
                                             t(batch["kps"]), input_size=cfg.input_size, target_sigma=cfg.sigma)
    np.testing.assert_allclose([float(losses[k]) for k in oc.LOSS_KEYS], g["loss_f32"], rtol=2e-6, atol=1e-9)
    gh = grads["heatmaps"].numpy()
    scale = np.abs(g["grad_hm_sub"]).max()
    assert np.abs(gh.reshape(-1)[::LARGE_STRIDE] - g["grad_hm_sub"]).max() <= 2e-6 * scale
    mom, wm = tile_moments(gh), g["grad_hm_moments"]
    assert (np.abs(mom[..., 1] - wm[..., 1]) / (wm[..., 1] + 1e-30)).max() <= 1e-5
    assert np.array_equal(grads["offsets"].numpy() != 0, g["grad_off"] != 0)
    c, s = oc.fusion_decode(t(batch["heatmaps"]), t(batch["offsets"]), float(g["alpha_param"]), float(g["fusion_weight"]))
    ok = (np.abs(g["dec_softargmax"] - np.floor(g["dec_softargmax"]) - 0.5) > 1e-3).all(-1)
    assert np.array_equal(s.numpy(), g["dec_scores"]) and np.abs(c.numpy() - g["dec_coords"])[ok].max() <= 2e-5
    kp, mv, idx = oc.decode_heatmaps(t(batch["heatmaps"]))
    assert np.array_equal(kp.numpy(), g["argmax_coords"]) and np.array_equal(mv.numpy(), g["argmax_vals"])
