"""GPU parity: the CUDA path (through torch ops -> ctypes -> the C ABI of
libgbcodec.so) against the CPU oracle and against the golden vectors produced by
the reference itself.  Tolerances are BASELINE.json's:

  integer peak indices        bit-exact
  encoded heatmaps            1e-6 relative (fp32)
  loss values and gradients   1e-5 relative
  decoded coordinates         1e-4 px
"""
import contextlib
import os

import numpy as np
import pytest
import torch

from oracle import heatmap_codec as oc
from tests import goldens, synth

pytestmark = pytest.mark.gpu

NAMES = list(synth.CONFIGS)
ENC_RTOL, LOSS_RTOL, COORD_ATOL = 1e-6, 1e-5, 1e-4


@pytest.fixture(scope="module")
def gb():
    import infantposeestimation_gaussianbias_b200 as pkg
    pkg.load()           # raises if libgbcodec.so is missing: no fallback
    from infantposeestimation_gaussianbias_b200 import ops
    return ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@contextlib.contextmanager
def tile_kernel():
    """The float32 entry point through the one-CTA-per-tile kernel (csrc/loss_tile.cu) instead of the persistent step
    kernel: the float16 and per-tile-mean entry points are instantiations of that kernel, and the tests that compare
    them with the float32 path BIT FOR BIT need the same arithmetic on both sides.  (The step kernel itself is checked
    against the goldens and the oracle like everything else.)"""
    old = os.environ.get("GBCODEC_STEP_KERNEL")
    os.environ["GBCODEC_STEP_KERNEL"] = "tile"
    try:
        yield
    finally:
        if old is None:
            del os.environ["GBCODEC_STEP_KERNEL"]
        else:
            os.environ["GBCODEC_STEP_KERNEL"] = old


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def grad_close(got, want, rel=LOSS_RTOL, what="", truth=None):
    """max-norm relative error: |got - want|_inf <= rel * |want|_inf, `want` being the
    reference's fp32 result.

    The offset-map gradient sits on the four bilinear taps around the soft-argmax
    coordinate, so it inherits that coordinate's fp32 noise (~1e-5 px on a 48..128 px
    axis, i.e. 1e-5..5e-5 relative on a tap weight): the reference's OWN fp32 result is
    1.2e-5 / 2.0e-5 / 4.7e-5 away from its fp64 result on the three golden batches.
    Where `truth` (the fp64 result) is given, a result that misses `rel` against the
    fp32 reference still passes if it is no further from the truth than 3x what the
    fp32 reference itself is."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = np.abs(want).max()
    err = np.abs(got - want).max()
    if err <= rel * scale + 1e-30:
        return
    if truth is not None:
        truth = np.asarray(truth, np.float64)
        ref_noise = np.abs(want - truth).max()
        mine = np.abs(got - truth).max()
        n_bad = int((np.abs(got - want) > rel * scale).sum())
        print(f"[parity] {what}: {n_bad} of {int((want != 0).sum())} non-zero entries miss {rel:g} against the float32 reference "
              f"({err / scale:.2e}); against float64: ours {mine / scale:.2e}, the float32 reference {ref_noise / scale:.2e}")
        # (not rare, and not a defect: at 1e-5 of the largest tap the float32 REFERENCE is itself 1.2e-5 .. 4.7e-5 away
        #  from its float64 evaluation — the tap weights inherit the soft-argmax's float32 noise — so what is bounded is
        #  the distance to the float64 truth, and the count is printed)
        assert mine <= max(rel * scale, 3 * ref_noise), \
            f"{what}: {mine / scale:.2e} from fp64 truth, the fp32 reference is {ref_noise / scale:.2e} from it"
        return
    assert False, f"{what}: max err {err:.3e} vs scale {scale:.3e} (ratio {err / max(scale, 1e-300):.2e})"


def half_integer_free(coords, eps=1e-3):
    """Tiles whose soft-argmax is within eps of a .5 boundary are excluded from the
    window-centre comparison (fp32 summation order decides the rounding)."""
    frac = np.abs(coords - np.floor(coords) - 0.5)
    return (frac > eps).all(axis=-1)


# ------------------------------------------------------------------ encode
@pytest.mark.parametrize("name", NAMES)
def test_encode(gb, name):
    cfg, batch, g = goldens.load(name)
    for kps, vis, want_t, want_w in ((batch["kps"], batch["vis"], g["target"], g["enc_weight"]),
                                     (g["edge_kps"], g["edge_vis"], g["edge_target"], g["edge_weight"])):
        target, weight = gb.encode(dev(kps), dev(vis), cfg.H, cfg.W, float(cfg.input_size[0]), float(cfg.input_size[1]), cfg.sigma)
        target, weight = target.cpu().numpy(), weight.cpu().numpy()
        assert np.array_equal(weight, want_w)
        assert np.array_equal(target != 0, want_t != 0), "support of the pasted patch differs"
        np.testing.assert_allclose(target, want_t, rtol=ENC_RTOL, atol=0)


def test_encode_large_batch_properties(gb):
    # BASELINE configs[1] size: peak value 1 at the integer pixel, tile sum bounded, weights follow the rule
    cfg = synth.CONFIGS["w32_256x192"]
    rng = np.random.default_rng(7)
    kps, vis = synth.make_keypoints(cfg, rng, 1024)
    target, weight = gb.encode(dev(kps), dev(vis), cfg.H, cfg.W, 192.0, 256.0, 2.0)
    ot, ow = oc.encode_targets(kps, vis, cfg.heatmap_size, cfg.input_size, 2.0)
    assert np.array_equal(weight.cpu().numpy(), ow)
    np.testing.assert_allclose(target.cpu().numpy(), ot, rtol=ENC_RTOL, atol=0)


# ------------------------------------------------------------------ argmax family
@pytest.mark.parametrize("name", NAMES)
def test_decode_heatmaps_bit_exact(gb, name):
    cfg, batch, g = goldens.load(name)
    hm = dev(batch["heatmaps"])
    c, v, idx = gb.decode_argmax(hm, 1)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), g["argmax_idx"])
    assert np.array_equal(v.cpu().numpy(), g["argmax_vals"])
    assert np.array_equal(c.cpu().numpy(), g["argmax_coords"])
    c0, _, _ = gb.decode_argmax(hm, 0)
    oc0, _, _ = oc.decode_heatmaps(t(batch["heatmaps"]), shift=False)
    assert np.array_equal(c0.cpu().numpy(), oc0.numpy())


def test_argmax_ties_and_borders(gb):
    h = torch.zeros(2, 3, 8, 12)
    h[0, 1, 2, 3] = 1.0; h[0, 1, 5, 1] = 1.0          # duplicate maximum: the first wins
    h[0, 2, 0, 5] = 2.0                                # border row: no shift
    h[1, 0, 4, 11] = 2.0                               # border column
    h[1, 1, 3, 4] = 1.0; h[1, 1, 3, 5] = 0.5; h[1, 1, 2, 4] = 0.25
    c, v, idx = gb.decode_argmax(h.cuda(), 1)
    oc_c, oc_v, oc_i = oc.decode_heatmaps(h, shift=True)
    assert np.array_equal(idx.cpu().numpy(), oc_i.numpy())
    assert np.array_equal(c.cpu().numpy(), oc_c.numpy())
    assert idx[0, 0].item() == 0 and idx[0, 1].item() == 2 * 12 + 3


# ------------------------------------------------------------------ decode
@pytest.mark.parametrize("name", NAMES)
def test_decode_against_golden(gb, name):
    cfg, batch, g = goldens.load(name)
    hm, off = dev(batch["heatmaps"]), dev(batch["offsets"])
    alpha = torch.tensor(float(g["alpha_param"])).cuda()
    fw = torch.tensor(float(g["fusion_weight"])).cuda()
    ok = half_integer_free(g["dec_softargmax"])
    print(f"[parity] {name} decode: {int((~ok).sum())} of {ok.size} tiles within 1e-3 px of a rounding boundary (H1) excluded")
    assert (~ok).sum() <= 0.1 * ok.size
    c, s, centre = gb.decode(hm, None, None, off, alpha, fw, 2, 3)
    assert np.array_equal(s.cpu().numpy(), g["dec_scores"])
    assert np.abs(c.cpu().numpy() - g["dec_coords"])[ok].max() <= COORD_ATOL
    want_centre = oc.window_centres(t(g["dec_softargmax"]), cfg.H, cfg.W).numpy()
    assert np.array_equal(centre.cpu().numpy()[ok], want_centre[ok])
    c, _, _ = gb.decode(hm, None, None, None, alpha, None, 2, 1)
    assert np.abs(c.cpu().numpy() - g["dec_coords_nooff"])[ok].max() <= COORD_ATOL
    c, _, _ = gb.decode(hm, None, None, None, None, None, 0, 0)
    assert np.abs(c.cpu().numpy() - g["dec_softargmax"]).max() <= COORD_ATOL
    # flip test
    cf, sf, _ = gb.decode(hm, dev(batch["heatmaps_flip"]), dev(batch["flip_perm"]), off, alpha, fw, 2, 3)
    avg = oc.flip_average(t(batch["heatmaps"]), t(batch["heatmaps_flip"]),
                          [p for p in oc.COCO_FLIP_PAIRS if p[0] < cfg.K and p[1] < cfg.K])
    okf = half_integer_free(oc.soft_argmax(avg)[0].numpy())
    assert np.array_equal(sf.cpu().numpy(), g["flip_scores"])
    assert np.abs(cf.cpu().numpy() - g["flip_coords"])[okf].max() <= COORD_ATOL


def test_decode_larger_batch_vs_oracle(gb):
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=3, B=64)
    hm, off = t(batch["heatmaps"]), t(batch["offsets"])
    want_c, want_s = oc.fusion_decode(hm, off, 0.5, 0.6224593312018546)
    soft = oc.soft_argmax(hm)[0]
    ok = half_integer_free(soft.numpy())
    c, s, centre = gb.decode(hm.cuda(), None, None, off.cuda(), torch.tensor(0.5).cuda(), torch.tensor(0.6224593312018546).cuda(), 2, 3)
    assert np.array_equal(s.cpu().numpy(), want_s.numpy())
    assert np.abs(c.cpu().numpy() - want_c.numpy())[ok].max() <= COORD_ATOL
    assert ok.mean() > 0.9        # flat tiles put the soft-argmax at (W-1)/2 = 23.5: ~5 % sit on the boundary
    # boundary tiles: the kernel's window centre must be one of the two neighbouring pixels, and with
    # THAT centre the oracle must reproduce the kernel's coordinates (so only the rounding is in doubt)
    centre = centre.cpu()
    assert (np.abs(centre.numpy() - soft.numpy()) <= 0.5 + 1e-3).all()
    alt_c, _ = oc.fusion_decode(hm, off, 0.5, 0.6224593312018546, centre=centre)
    assert np.abs(c.cpu().numpy() - alt_c.numpy()).max() <= COORD_ATOL


@pytest.mark.parametrize("name,radius", [("w32_256x192", 2), ("w32_256x192", 4), ("w32_256x192", 0)])
def test_warp_per_tile_kernels_equal_the_cta_per_tile_kernels(gb, name, radius, monkeypatch):
    """The warp-per-tile decode / arg-max kernels (the default for tiles of at most 16 KB) against the CTA-per-tile ones on
    the same tiles: arg-max family bit for bit; decode scores and window centres equal, coordinates to 2e-5 px (the
    softmax sums run in a different order)."""
    cfg = synth.CONFIGS[name]
    batch = synth.make_batch(cfg, seed=21, B=37)                      # 629 tiles: not a multiple of the warps per CTA
    hm, off = dev(batch["heatmaps"]), dev(batch["offsets"])
    alpha, fw = torch.tensor(0.3).cuda(), torch.tensor(0.7).cuda()
    flags = (1 if radius > 0 else 0) | 2
    got = [x.cpu().numpy() for x in gb.decode(hm, None, None, off, alpha, fw, radius, flags)]
    got_a = [[x.cpu().numpy() for x in gb.decode_argmax(hm, mode)] for mode in (0, 1, 2)]
    monkeypatch.setenv("GBCODEC_DECODE_KERNEL", "tile")
    monkeypatch.setenv("GBCODEC_ARGMAX_KERNEL", "tile")
    want = [x.cpu().numpy() for x in gb.decode(hm, None, None, off, alpha, fw, radius, flags)]
    want_a = [[x.cpu().numpy() for x in gb.decode_argmax(hm, mode)] for mode in (0, 1, 2)]
    for g3, w3 in zip(got_a, want_a):
        for g, w in zip(g3, w3):
            assert np.array_equal(g, w)
    assert np.array_equal(got[1], want[1])
    same = (got[2] == want[2]).all(axis=-1)
    print(f"[parity] warp-per-tile decode r={radius}: {int((~same).sum())} of {same.size} window centres differ (rounding boundary)")
    assert same.mean() >= 0.99
    assert np.abs(got[0] - want[0])[same].max() <= 2e-5


# ------------------------------------------------------------------ loss
def run_loss(gb, cfg, batch, *, on_the_fly, with_grads=True, utw=True, denoms=None, grad_scale=None, decode=False,
             lambdas=oc.DEFAULT_LAMBDAS):
    pairs = [v for p in oc.skeleton_for(cfg.K) for v in p]
    return gb.fusion_loss(
        dev(batch["heatmaps"]), dev(batch["offsets"]), dev(batch["variances"]),
        None if on_the_fly else dev(batch["target"]),
        dev(batch["weight"] if not on_the_fly else batch["vis"][..., None]), dev(batch["kps"]),
        denoms, grad_scale, float(cfg.input_size[0]), float(cfg.input_size[1]), list(lambdas),
        cfg.sigma, cfg.sigma, utw, pairs, with_grads, decode,
        torch.tensor(0.5).cuda() if decode else None, torch.tensor(0.6224593312018546).cuda() if decode else None, 2, 3)[:6]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("on_the_fly", [False, True])
def test_loss_and_grads_against_golden(gb, name, on_the_fly):
    cfg, batch, g = goldens.load(name)
    losses, ghm, goff, gvar, _, _ = run_loss(gb, cfg, batch, on_the_fly=on_the_fly)
    np.testing.assert_allclose(losses.cpu().numpy(), g["loss_f32"], rtol=LOSS_RTOL, atol=1e-9)
    np.testing.assert_allclose(losses.cpu().numpy(), g["loss_f64"], rtol=LOSS_RTOL, atol=1e-9)
    grad_close(ghm.cpu().numpy(), g["grad_hm"], what="grad heatmaps")
    go = goff.cpu().numpy()
    assert np.array_equal(go != 0, g["grad_off"] != 0)
    grad_close(go, g["grad_off"], what="grad offsets", truth=g["grad_off_f64"])
    gv = gvar.cpu().numpy()
    grad_close(gv, np.broadcast_to(g["grad_var_tile"][:, :, None, None], gv.shape), what="grad variances")


@pytest.mark.parametrize("utw", [True, False])
def test_loss_vs_oracle_batch64(gb, utw):
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=11, B=64)
    want_l, want_g = oc.fusion_loss_and_grads(t(batch["heatmaps"]), t(batch["offsets"]), t(batch["variances"]),
                                              t(batch["target"]), t(batch["weight"]), t(batch["kps"]),
                                              input_size=cfg.input_size, use_target_weight=utw)
    d = lambda k: t(batch[k]).double()
    _, truth_g = oc.fusion_loss_and_grads(d("heatmaps"), d("offsets"), d("variances"), d("target"), d("weight"), d("kps"),
                                          input_size=cfg.input_size, use_target_weight=utw)
    losses, ghm, goff, gvar, _, _ = run_loss(gb, cfg, batch, on_the_fly=True, utw=utw)
    got = losses.cpu().numpy()
    want = np.array([float(want_l[k]) for k in oc.LOSS_KEYS])
    np.testing.assert_allclose(got, want, rtol=LOSS_RTOL, atol=1e-9)
    grad_close(ghm.cpu().numpy(), want_g["heatmaps"].numpy(), what="grad heatmaps")
    grad_close(goff.cpu().numpy(), want_g["offsets"].numpy(), what="grad offsets", truth=truth_g["offsets"].numpy())
    grad_close(gvar.cpu().numpy(), want_g["variances"].numpy(), what="grad variances")


def test_loss_forward_only_and_step(gb):
    cfg, batch, g = goldens.load("w32_256x192")
    losses, ghm, goff, gvar, _, _ = run_loss(gb, cfg, batch, on_the_fly=True, with_grads=False)
    assert ghm.numel() == 0 and goff.numel() == 0 and gvar.numel() == 0
    np.testing.assert_allclose(losses.cpu().numpy(), g["loss_f32"], rtol=LOSS_RTOL, atol=1e-9)
    # fused step: same losses and gradients plus the decode of the same tiles
    l2, ghm2, goff2, gvar2, coords, scores = run_loss(gb, cfg, batch, on_the_fly=True, decode=True)
    np.testing.assert_allclose(l2.cpu().numpy(), g["loss_f32"], rtol=LOSS_RTOL, atol=1e-9)
    grad_close(ghm2.cpu().numpy(), g["grad_hm"], what="step grad heatmaps")
    ok = half_integer_free(g["dec_softargmax"])
    assert np.abs(coords.cpu().numpy() - g["dec_coords"])[ok].max() <= COORD_ATOL
    assert np.array_equal(scores.cpu().numpy(), g["dec_scores"])


def test_loss_sharded_denominators(gb):
    """Two shards with the global denominators reproduce the full-batch loss and gradients."""
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=5, B=8)
    full_l, full_ghm, _, _, _, _ = run_loss(gb, cfg, batch, on_the_fly=True)
    pairs = [v for p in oc.skeleton_for(cfg.K) for v in p]
    den = gb.loss_denominators(dev(batch["vis"]), dev(batch["kps"]), False, cfg.H, cfg.W, 192.0, 256.0, cfg.sigma, pairs)
    want_den = oc.loss_denominators(t(batch["weight"]), cfg.K)
    np.testing.assert_allclose(den.cpu().numpy(), [float(want_den[0]), float(want_den[1])], rtol=1e-6)
    acc = torch.zeros(7, device="cuda")
    parts = []
    for sl in (slice(0, 3), slice(3, 8)):
        sub = {k: v[sl] for k, v in batch.items() if k != "flip_perm"}
        l, gh, _, _, _, _ = run_loss(gb, cfg, sub, on_the_fly=True, denoms=den)
        acc += l
        parts.append(gh)
    np.testing.assert_allclose(acc.cpu().numpy(), full_l.cpu().numpy(), rtol=LOSS_RTOL)
    grad_close(torch.cat(parts).cpu().numpy(), full_ghm.cpu().numpy(), what="sharded grad")


def test_module_autograd_and_scaling(gb):
    """FusionPoseLoss drop-in: dict of 7, backward through total_loss, a scaled loss
    (GradScaler-style), and a per-term upstream gradient (recompute path)."""
    from infantposeestimation_gaussianbias_b200 import FusionPoseLoss
    cfg, batch, g = goldens.load("w32_256x192")
    loss_fn = FusionPoseLoss(target_sigma=cfg.sigma)

    def fresh():
        o = {"heatmaps": dev(batch["heatmaps"]).requires_grad_(True), "offsets": dev(batch["offsets"]).requires_grad_(True),
             "variances": dev(batch["variances"]).requires_grad_(True)}
        return o

    o = fresh()
    out = loss_fn(o, dev(batch["target"]), dev(batch["weight"]), dev(batch["kps"]), input_size=cfg.input_size, heatmap_size=(cfg.H, cfg.W))
    assert list(out) == list(oc.LOSS_KEYS) and all(v.dim() == 0 for v in out.values())
    out["total_loss"].backward()
    grad_close(o["heatmaps"].grad.cpu().numpy(), g["grad_hm"], what="module grad")
    grad_close(o["offsets"].grad.cpu().numpy(), g["grad_off"], what="module grad off", truth=g["grad_off_f64"])

    o = fresh()   # scaled loss: rescale path
    out = loss_fn(o, None, dev(batch["vis"][..., None]), dev(batch["kps"]), input_size=cfg.input_size)
    (out["total_loss"] * 1024.0).backward()
    grad_close(o["heatmaps"].grad.cpu().numpy() / 1024.0, g["grad_hm"], what="scaled grad")

    o = fresh()   # known scale up front: nothing to do in backward
    out = loss_fn(o, None, dev(batch["vis"][..., None]), dev(batch["kps"]), input_size=cfg.input_size,
                  grad_scale=torch.tensor(1024.0).cuda())
    (out["total_loss"] * 1024.0).backward()
    grad_close(o["heatmaps"].grad.cpu().numpy() / 1024.0, g["grad_hm"], what="pre-scaled grad")

    o = fresh()   # only two of the six terms: per-term recompute
    out = loss_fn(o, dev(batch["target"]), dev(batch["weight"]), dev(batch["kps"]), input_size=cfg.input_size)
    (out["heatmap_loss"] + 3.0 * out["shape_loss"]).backward()
    ref = {k: t(batch[k]).clone().requires_grad_(True) for k in ("heatmaps", "offsets", "variances")}
    rl = oc.fusion_loss(ref["heatmaps"], ref["offsets"], ref["variances"], t(batch["target"]), t(batch["weight"]), t(batch["kps"]),
                        input_size=cfg.input_size, target_sigma=cfg.sigma)
    (rl["heatmap_loss"] + 3.0 * rl["shape_loss"]).backward()
    grad_close(o["heatmaps"].grad.cpu().numpy(), ref["heatmaps"].grad.numpy(), what="per-term grad")
    assert float(o["offsets"].grad.abs().max()) == 0.0


@pytest.mark.parametrize("half", [False, True])
def test_backward_twice_on_one_graph(gb, half):
    """the backward adjusts the stored gradients in place (rescale / per-term recompute), so a second backward
    through the same graph must compare the upstream it gets with what the stored gradients hold NOW, not with what the
    forward assumed.  Sequence on ONE graph: heatmap_loss alone (recompute) -> total_loss (would be "nothing to do"
    against the forward's assumption) -> 3 * total_loss (rescale) -> 3 * total_loss again (nothing to do); each against
    the oracle's autograd for the same upstream."""
    from infantposeestimation_gaussianbias_b200 import FusionPoseLoss
    cfg, batch, g = goldens.load("w32_256x192")
    loss_fn = FusionPoseLoss(target_sigma=cfg.sigma)
    cast = (lambda x: x.half()) if half else (lambda x: x)
    base = {k: cast(dev(batch[k])) for k in ("heatmaps", "offsets", "variances")}
    o = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    ref_in = {k: v.detach().float().cpu().clone().requires_grad_(True) for k, v in base.items()}
    rl = oc.fusion_loss(ref_in["heatmaps"], ref_in["offsets"], ref_in["variances"], t(batch["target"]), t(batch["weight"]), t(batch["kps"]),
                        input_size=cfg.input_size, target_sigma=cfg.sigma)
    out = loss_fn(o, dev(batch["target"]), dev(batch["weight"]), dev(batch["kps"]), input_size=cfg.input_size)
    assert not out["total_loss"].requires_grad or out["total_loss"].grad_fn is not None
    steps = [("heatmap_loss alone", lambda d: d["heatmap_loss"]), ("total_loss", lambda d: d["total_loss"]),
             ("3 x total_loss", lambda d: 3.0 * d["total_loss"]), ("3 x total_loss again", lambda d: 3.0 * d["total_loss"]),
             ("shape + total", lambda d: d["shape_loss"] + d["total_loss"])]
    for what, pick in steps:
        for v in o.values():
            v.grad = None
        for v in ref_in.values():
            v.grad = None
        pick(out).backward(retain_graph=True)
        pick(rl).backward(retain_graph=True)
        for k in ("heatmaps", "variances"):
            got = o[k].grad.float().cpu().numpy()
            want = np.zeros_like(got) if ref_in[k].grad is None else ref_in[k].grad.numpy()     # a term that does not see this map
            tol = 2e-3 if half else LOSS_RTOL              # half: the gradient is rounded to float16 once
            # half: gradients of ~1e-6 (no loss scale here) sit in float16's subnormal range, spacing 6e-8
            assert np.abs(got - want).max() <= tol * max(np.abs(want).max(), 1e-30) + (6e-8 if half else 1e-12), (what, k)
        go = o["offsets"].grad.float().cpu().numpy()
        wo = ref_in["offsets"].grad
        wo = np.zeros_like(go) if wo is None else wo.numpy()
        assert np.abs(go - wo).max() <= (2e-3 if half else 3e-5) * max(np.abs(wo).max(), 1e-30) + (6e-8 if half else 1e-12), (what, "offsets")


def test_coords_and_scores_are_not_differentiable(gb):
    """only the seven losses carry a graph; decode outputs and the stored gradients are plain tensors."""
    from infantposeestimation_gaussianbias_b200 import FusionPoseLoss
    cfg, batch, g = goldens.load("w32_256x192")
    o = {k: dev(batch[k]).requires_grad_(True) for k in ("heatmaps", "offsets", "variances")}
    out = FusionPoseLoss(target_sigma=cfg.sigma)(o, None, dev(batch["vis"][..., None]), dev(batch["kps"]), input_size=cfg.input_size,
                                                 decode={"alpha_param": torch.tensor(0.5).cuda(), "fusion_weight": torch.tensor(0.62).cuda()})
    assert out["total_loss"].requires_grad and not out["coords"].requires_grad and not out["scores"].requires_grad
    out["coords"].cpu().numpy()          # would raise on a tensor that requires grad


def test_identical_tiles_tie_rule(gb):
    """All channels equal (a zero-initialised last layer): every min() ties, ATen splits the
    gradient evenly; the kernel must reproduce that."""
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=2, B=2)
    batch["heatmaps"][:] = batch["heatmaps"][:, :1]
    want_l, want_g = oc.fusion_loss_and_grads(t(batch["heatmaps"]), t(batch["offsets"]), t(batch["variances"]),
                                              t(batch["target"]), t(batch["weight"]), t(batch["kps"]), input_size=cfg.input_size)
    losses, ghm, _, _, _, _ = run_loss(gb, cfg, batch, on_the_fly=False)
    np.testing.assert_allclose(losses.cpu().numpy(), [float(want_l[k]) for k in oc.LOSS_KEYS], rtol=LOSS_RTOL, atol=1e-9)
    grad_close(ghm.cpu().numpy(), want_g["heatmaps"].numpy(), what="tie grad")


# ------------------------------------------------------------------ the large golden set (BASELINE configs[0] size)
def _report(tag, **kw):
    print(f"[parity] {tag}: " + ", ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in kw.items()))


@pytest.mark.parametrize("name", NAMES)
def test_fused_step_against_large_goldens(gb, name):
    """Reference-generated goldens at BASELINE configs[0] size (B = 32 at 64x48; B = 8 for 96x72 and 128x128): losses,
    heatmap gradient ELEMENT-WISE (every 7th element, |err| <= 1e-5 |want| + 1e-5 max|g| of its tile) and through six
    moments of every tile, offset gradient (support bit-equal; values against the float32 reference AND its float64
    evaluation: at 1e-5 of the largest tap the float32 reference is itself 1e-5 .. 5e-5 from float64, so the bound is on
    the distance to float64 — ours at most 3x the reference's — and the number of entries beyond 1e-5 is printed),
    variance gradient, decode.  Everything that is excluded is counted, printed and bounded."""
    from tests.golden.make_golden import LARGE_STRIDE, tile_moments
    cfg, batch, g = goldens.load_large(name)
    pairs = [v for p in oc.skeleton_for(cfg.K) for v in p]
    alpha, fw = torch.tensor(float(g["alpha_param"])).cuda(), torch.tensor(float(g["fusion_weight"])).cuda()
    for on_the_fly in (False, True):
        res = gb.fusion_loss(dev(batch["heatmaps"]), dev(batch["offsets"]), dev(batch["variances"]),
                             None if on_the_fly else dev(g["target"]), dev(batch["vis"] if on_the_fly else g["enc_weight"]),
                             dev(batch["kps"]), None, None, float(cfg.input_size[0]), float(cfg.input_size[1]), list(oc.DEFAULT_LAMBDAS),
                             cfg.sigma, cfg.sigma, True, pairs, True, True, alpha, fw, 2, 3)
        losses, ghm, goff, gvar, coords, scores = [x.cpu().numpy() for x in res[:6]]
        loss_err = np.abs(losses - g["loss_f32"]) / np.maximum(np.abs(g["loss_f32"]), 1e-12)
        np.testing.assert_allclose(losses, g["loss_f32"], rtol=LOSS_RTOL, atol=1e-9)
        # heatmap gradient, element-wise on the stored subset
        sub = ghm.reshape(-1)[::LARGE_STRIDE].astype(np.float64)
        want = g["grad_hm_sub"].astype(np.float64)
        tile_of = (np.arange(sub.size) * LARGE_STRIDE) // (cfg.H * cfg.W)
        floor = g["grad_hm_moments"][..., 3].reshape(-1)[tile_of]
        excess = np.abs(sub - want) / (LOSS_RTOL * np.abs(want) + LOSS_RTOL * floor + 1e-30)
        assert excess.max() <= 1.0, f"{name}: element-wise gradient error {excess.max():.2f}x the bound"
        # ... and every pixel of every tile through the tile's moments (scale: sum |g| of the tile, times the axis length for the first moments)
        mom, wm = tile_moments(ghm), g["grad_hm_moments"]
        l1 = wm[..., 1] + 1e-30
        mom_err = max(np.abs(mom[..., 0] - wm[..., 0]).max() / l1.max(), (np.abs(mom[..., 1] - wm[..., 1]) / l1).max(),
                      (np.abs(mom[..., 4] - wm[..., 4]) / (l1 * cfg.W)).max(), (np.abs(mom[..., 5] - wm[..., 5]) / (l1 * cfg.H)).max(),
                      (np.abs(mom[..., 3] - wm[..., 3]) / (wm[..., 3] + 1e-30)).max())
        assert mom_err <= 2e-5, f"{name}: tile moments off by {mom_err:.2e}"
        # offset gradient: same support; values within 1e-5 max-norm, or (counted) no further from float64 than 3x the float32 reference
        assert np.array_equal(goff != 0, g["grad_off"] != 0)
        scale = np.abs(g["grad_off"]).max()
        bad = np.abs(goff - g["grad_off"]) > LOSS_RTOL * scale
        ref_noise = np.abs(g["grad_off"] - g["grad_off_f64"]).max()
        escaped = int(bad.sum())
        if escaped:
            assert np.abs(goff - g["grad_off_f64"])[bad].max() <= max(LOSS_RTOL * scale, 3 * ref_noise)
        nz = int((g["grad_off"] != 0).sum())
        ours_vs_f64 = np.abs(goff - g["grad_off_f64"]).max()
        assert ours_vs_f64 <= max(LOSS_RTOL * scale, 3 * ref_noise)
        np.testing.assert_allclose(gvar[:, :, 0, 0], g["grad_var_tile"], rtol=1e-4, atol=1e-12)
        assert np.all(gvar == gvar[:, :, :1, :1])
        # decode
        ok = half_integer_free(g["dec_softargmax"])
        excluded = int((~ok).sum())
        assert excluded <= 0.1 * ok.size
        assert np.array_equal(scores, g["dec_scores"])
        cerr = np.abs(coords - g["dec_coords"]).max(-1)
        assert cerr[ok].max() <= COORD_ATOL
        _report(f"{name} large golden ({'on-the-fly' if on_the_fly else 'given'} target)", tiles=ok.size, loss_rel_err=float(loss_err.max()),
                grad_hm_elementwise_worst_over_bound=float(excess.max()), grad_hm_moment_err=float(mom_err),
                grad_off_beyond_1e5_of_f32_ref=escaped, grad_off_nonzeros=nz, grad_off_ours_vs_f64=float(ours_vs_f64 / scale),
                grad_off_ref_f32_vs_f64=float(ref_noise / scale),
                decode_h1_excluded=excluded, decode_max_err_px=float(cerr[ok].max()))
    # arg-max family: integer indices bit-exact
    c, v, idx = gb.decode_argmax(dev(batch["heatmaps"]), 1)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), g["argmax_idx"]) and np.array_equal(v.cpu().numpy(), g["argmax_vals"])
    assert np.array_equal(c.cpu().numpy(), g["argmax_coords"])
    cf, sf, _ = gb.decode(dev(batch["heatmaps"]), dev(batch["heatmaps_flip"]), dev(batch["flip_perm"]), dev(batch["offsets"]), alpha, fw, 2, 3)
    avg = oc.flip_average(t(batch["heatmaps"]), t(batch["heatmaps_flip"]), [p for p in oc.COCO_FLIP_PAIRS if p[0] < cfg.K and p[1] < cfg.K])
    okf = half_integer_free(oc.soft_argmax(avg)[0].numpy())
    assert (~okf).sum() <= 0.1 * okf.size and np.array_equal(sf.cpu().numpy(), g["flip_scores"])
    assert np.abs(cf.cpu().numpy() - g["flip_coords"])[okf].max() <= COORD_ATOL


# ------------------------------------------------------------------ next-row decoders (utils/postprocess.py)
def test_postprocess_family(gb):
    from infantposeestimation_gaussianbias_b200 import postprocess as pp
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=9, B=4)
    hm = np.abs(batch["heatmaps"]) + 0.01          # the linear-weight centroid assumes positive maps
    c, v = pp.get_max_preds(dev(hm))
    oc_c, oc_v, _ = oc.decode_heatmaps(t(hm), shift=False)
    assert np.array_equal(c.cpu().numpy(), oc_c.numpy()) and np.array_equal(v.cpu().numpy()[..., 0], oc_v.numpy())


# ------------------------------------------------------------------ errors, no fallback
def test_errors_are_loud(gb):
    from infantposeestimation_gaussianbias_b200 import GbcodecError
    with pytest.raises(RuntimeError):
        gb.decode_argmax(torch.zeros(1, 1, 8, 8), 0)              # CPU tensor
    with pytest.raises(GbcodecError):
        gb.decode_argmax(torch.zeros(1, 1, 8, 6).cuda(), 0)       # W % 4 != 0
    with pytest.raises(GbcodecError):
        gb.decode_argmax(torch.zeros(1, 1, 8, 8).cuda(), 7)       # bad mode
    with pytest.raises(GbcodecError):
        gb.decode(torch.zeros(1, 1, 8, 8).cuda(), None, None, None, None, None, 2, 1)   # REFINE without alpha


# ---------------------------------------------------------------------- host-buffer step
@pytest.mark.parametrize("stage_offsets", [False, True])
def test_host_buffer_step_matches_resident_step(gb, stage_offsets):
    """HostCodecStep (pinned host buffers in, chunked H2D pipeline, offsets read in place over
    PCIe or staged) must give what one resident pass over the whole batch gives."""
    from infantposeestimation_gaussianbias_b200 import _native as N
    from infantposeestimation_gaussianbias_b200.host_step import HostCodecStep
    cfg = synth.CONFIGS["w32_256x192"]
    B = 24
    batch = synth.make_batch(cfg, seed=5, B=B)
    K, (W, H) = cfg.K, cfg.heatmap_size
    pin = lambda k: t(batch[k]).pin_memory()
    hs = HostCodecStep(B, K, H, W, cfg.input_size, cfg.sigma, chunk_images=7, stage_offsets=stage_offsets)
    out = hs(pin("heatmaps"), pin("offsets"), pin("variances"), pin("kps"), pin("vis"))
    pairs = gb.pairs_flat(oc.COCO_SKELETON)
    alpha, fw = torch.tensor([0.5]).cuda(), torch.tensor([0.6224593312018546]).cuda()
    res = gb.fusion_loss(dev(batch["heatmaps"]), dev(batch["offsets"]), dev(batch["variances"]), None, dev(batch["vis"]),
                         dev(batch["kps"]), None, None, float(cfg.input_size[0]), float(cfg.input_size[1]),
                         list(oc.DEFAULT_LAMBDAS), cfg.sigma, cfg.sigma, True, pairs, True, True, alpha, fw, 2,
                         N.DECODE_REFINE | N.DECODE_APPLY_OFFSET)
    np.testing.assert_allclose(out["losses"].numpy(), res[0].cpu().numpy(), rtol=2e-6)
    assert np.array_equal(out["coords"].numpy(), res[4].cpu().numpy())
    assert np.array_equal(out["scores"].numpy(), res[5].cpu().numpy())
    # EVERY chunk's gradients are kept, in persistent (B, ...) tensors: the resident step's gradients of the whole batch
    for got, want in zip(hs.grads, res[1:4]):
        assert got.shape == want.shape and torch.equal(got, want)
    assert hs.launches == 2 + 3 * ((B + 6) // 7)          # counted by the library: 2 for the normalisers, 3 kernels per chunk (memsets are not kernels)
    assert hs.h2d_bytes < (2 if not stage_offsets else 4) * B * K * H * W * 4 * 1.1


@pytest.mark.parametrize("flip", [False, True])
def test_host_buffer_decode_matches_resident_decode(gb, flip):
    """HostDecode (pinned heatmaps in, chunked H2D pipeline, offsets read in place over PCIe) gives the resident
    decode's coordinates and scores bit for bit, and those agree with the oracle (1e-4 px away from H1 tiles)."""
    from infantposeestimation_gaussianbias_b200 import _native as N
    from infantposeestimation_gaussianbias_b200.host_step import HostDecode
    from infantposeestimation_gaussianbias_b200.pose_estimator import flip_permutation
    cfg = synth.CONFIGS["hrformer_384x288"]
    B = 19
    batch = synth.make_batch(cfg, seed=6, B=B)
    K, (W, H) = cfg.K, cfg.heatmap_size
    pin = lambda k: t(batch[k]).pin_memory()
    hd = HostDecode(B, K, H, W, flip=flip, chunk_images=5, flip_pairs=oc.COCO_FLIP_PAIRS)
    out = hd(pin("heatmaps"), pin("heatmaps_flip") if flip else None, pin("offsets"))
    perm = flip_permutation(K, oc.COCO_FLIP_PAIRS, torch.device("cuda")) if flip else None
    alpha, fw = torch.tensor([0.5]).cuda(), torch.tensor([0.6224593312018546]).cuda()
    c, s, _ = gb.decode(dev(batch["heatmaps"]), dev(batch["heatmaps_flip"]) if flip else None, perm, dev(batch["offsets"]),
                        alpha, fw, 2, N.DECODE_REFINE | N.DECODE_APPLY_OFFSET)
    assert np.array_equal(out["coords"].numpy(), c.cpu().numpy())
    assert np.array_equal(out["scores"].numpy(), s.cpu().numpy())
    hm = t(batch["heatmaps"])
    flipped = t(batch["heatmaps_flip"]) if flip else None
    want_c, want_s = oc.fusion_decode(hm, t(batch["offsets"]), 0.5, 0.6224593312018546, True, True, 2, heatmaps_of_flipped_input=flipped)
    coarse, _ = oc.soft_argmax(oc.flip_average(hm, flipped) if flip else hm)
    safe = ((coarse - torch.floor(coarse) - 0.5).abs() > 1e-3).all(dim=-1).numpy()          # H1: away from the rounding boundary
    assert safe.mean() > 0.9
    np.testing.assert_allclose(out["coords"].numpy()[safe], want_c.numpy()[safe], atol=1e-4, rtol=0)
    np.testing.assert_allclose(out["scores"].numpy(), want_s.numpy(), rtol=1e-6 if flip else 0, atol=0)
    assert hd.h2d_bytes < (2 if flip else 1) * B * K * H * W * 4 * 1.1 and hd.launches == 4


# ---------------------------------------------------------------------- AMP: fp16 head outputs (train.py:171)
def same_up_to_tie_pixels(g16, g32, what):
    # 1024 = 2^10: scaling commutes with every rounding, so the two paths agree bit for bit — except on pixels whose
    # logit equals a limb partner's exactly (frequent in half precision): their share of the overlap gradient is
    # added in a second store, i.e. rounded twice (<= 1 ulp of half)
    want = g32.half()
    frac = (g16 == want).float().mean().item()
    assert frac > 0.9999, (what, frac)
    ulp = torch.maximum(want.float().abs(), torch.tensor(6.1e-5, device=want.device)) * 2.0 ** -10
    assert bool(((g16.float() - want.float()).abs() <= ulp).all()), what


@pytest.mark.parametrize("name", NAMES)
def test_fp16_head_outputs_under_autocast(gb, name):
    """Under autocast the head's outputs reach the loss in fp16.  The float16 kernels up-cast the values where they
    enter, sum in fp32 (what ATen's softmax / mse do under autocast) and round the gradients to fp16 once, after the
    upstream factor (the loss scale) has been applied: losses must equal the fp32 path's on the up-cast inputs bit for
    bit, gradients must equal that path's gradients rounded to fp16; the decode must equal the fp32 decode."""
    from infantposeestimation_gaussianbias_b200 import FusionPoseLoss, decode_outputs
    cfg = synth.CONFIGS[name]
    batch = synth.make_batch(cfg, seed=23, B=4)
    half = {k: dev(batch[k]).half() for k in ("heatmaps", "offsets", "variances")}
    loss_fn = FusionPoseLoss(target_sigma=cfg.sigma)
    assert loss_fn._half_maps(half)           # the float16 kernels take these shapes: nothing is up-cast in HBM
    with torch.autocast("cuda", dtype=torch.float16):
        o16 = {k: v.clone().requires_grad_(True) for k, v in half.items()}
        out16 = loss_fn(o16, dev(batch["target"]), dev(batch["weight"]), dev(batch["kps"]), input_size=cfg.input_size)
        scale = torch.tensor(1024.0).cuda()                   # a GradScaler-style loss scale
        (out16["total_loss"] * scale).backward()
    o32 = {k: v.float().requires_grad_(True) for k, v in half.items()}
    with tile_kernel():
        out32 = loss_fn(o32, dev(batch["target"]), dev(batch["weight"]), dev(batch["kps"]), input_size=cfg.input_size)
        (out32["total_loss"] * 1024.0).backward()
    # the same with targets built in the kernel, the decode fused in and per-term upstream gradients
    dec = {"alpha_param": torch.tensor(0.5).cuda(), "fusion_weight": torch.tensor(0.62).cuda()}
    mix = lambda o: 2048.0 * o["total_loss"] + 512.0 * o["heatmap_loss"] - 256.0 * o["shape_loss"]
    p16 = {k: v.clone().requires_grad_(True) for k, v in half.items()}
    q16 = loss_fn(p16, None, dev(batch["vis"]), dev(batch["kps"]), input_size=cfg.input_size, decode=dec)
    mix(q16).backward()
    p32 = {k: v.float().requires_grad_(True) for k, v in half.items()}
    # (no tile_kernel() here: with targets built in the kernel both element types take the same kernel of their shape —
    # the persistent step kernel for 64x48, the one-CTA-per-tile kernel otherwise — and the per-term upstream sends
    # both backwards through the recompute of the one-CTA-per-tile kernel)
    q32 = loss_fn(p32, None, dev(batch["vis"]), dev(batch["kps"]), input_size=cfg.input_size, decode=dec)
    mix(q32).backward()
    assert torch.equal(q16["coords"], q32["coords"]) and torch.equal(q16["scores"], q32["scores"])
    for k in oc.LOSS_KEYS:
        assert float(q16[k]) == float(q32[k])
    for k in p16:
        assert p16[k].grad.dtype == torch.float16
        # per-term weights go through the recompute path in both: lam * upstream is formed the same way
        same_up_to_tie_pixels(p16[k].grad, p32[k].grad, f"{k}, per-term upstream")
    for k in oc.LOSS_KEYS:
        assert out16[k].dtype == torch.float32 and float(out16[k]) == float(out32[k])
    for k in o16:
        assert o16[k].grad.dtype == torch.float16
        same_up_to_tie_pixels(o16[k].grad, o32[k].grad, k)
    # second step with the same module: the expected upstream (1024, remembered on the device) now holds, so the fused
    # pass's stored gradients are the result; a third step with another scale falls back to computing them again
    for scale_now in (1024.0, 64.0):
        r16 = {k: v.clone().requires_grad_(True) for k, v in half.items()}
        with torch.autocast("cuda", dtype=torch.float16):
            (loss_fn(r16, dev(batch["target"]), dev(batch["weight"]), dev(batch["kps"]), input_size=cfg.input_size)["total_loss"] * scale_now).backward()
        for k in r16:
            same_up_to_tie_pixels(r16[k].grad, o32[k].grad * (scale_now / 1024.0), f"{k} at scale {scale_now}")
    a = torch.tensor(0.5).cuda()
    fw = torch.sigmoid(torch.tensor(0.5)).cuda()
    c16, s16 = decode_outputs({**half, "fusion_weight": fw}, a)
    c32, s32 = decode_outputs({**{k: v.float() for k, v in half.items()}, "fusion_weight": fw}, a)
    assert torch.equal(c16, c32) and torch.equal(s16, s32) and c16.dtype == torch.float32


@contextlib.contextmanager
def env(name, value):
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        yield
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


def test_fp16_step_kernel_many_tiles_per_cta(gb):
    """The float16 instantiation of the persistent step kernel (64x48, targets built in the kernel; bulk copies of raw
    halves, a four-deep partner ring) on a batch that gives every CTA several tiles, weight-0 tiles and tiles with up to
    four limb partners: against the float32 instantiation on the up-cast maps the seven losses, coordinates and scores are
    BIT-equal and the gradients are that path's gradients rounded to half (tie pixels: rounded twice); against the
    one-CTA-per-tile float16 kernel (GBCODEC_STEP_F16=tile) everything agrees to float32 round-off."""
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=77, B=96)
    half = {k: dev(batch[k]).half() for k in ("heatmaps", "offsets", "variances")}
    vis, kps = dev(batch["vis"]), dev(batch["kps"])
    alpha, fw = torch.tensor(0.5).cuda(), torch.tensor(0.62).cuda()
    scale = torch.tensor([1024.0]).cuda()
    from infantposeestimation_gaussianbias_b200 import FusionPoseLoss, ops
    loss_fn = FusionPoseLoss(target_sigma=cfg.sigma)
    sk = ops.pairs_flat(loss_fn.pairs_for(cfg.K))
    DF = 1 | 2

    def f16():
        return ops.fusion_loss_f16(half["heatmaps"], half["offsets"], half["variances"], None, vis, kps, None, scale,
                                   float(cfg.input_size[0]), float(cfg.input_size[1]), loss_fn.lambdas, float(cfg.sigma), float(cfg.sigma), True, sk,
                                   True, True, alpha, fw, 2, DF)

    def f32():
        return ops.fusion_loss(half["heatmaps"].float(), half["offsets"].float(), half["variances"].float(), None, vis, kps, None, scale,
                               float(cfg.input_size[0]), float(cfg.input_size[1]), loss_fn.lambdas, float(cfg.sigma), float(cfg.sigma), True, sk,
                               True, True, alpha, fw, 2, DF, 0, False)

    a = f16()
    b = f32()
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]), (a[0], b[0])                       # losses7
    assert torch.equal(a[1], b[4]) and torch.equal(a[2], b[5])          # coords, scores
    for i, k in ((3, "heatmaps"), (4, "offsets"), (5, "variances")):
        assert a[i].dtype == torch.float16
        same_up_to_tie_pixels(a[i], b[i - 2], f"{k} (step kernel, float16 against float32)")
    with env("GBCODEC_STEP_F16", "tile"):
        c = f16()
    torch.cuda.synchronize()
    np.testing.assert_allclose(a[0].cpu().numpy(), c[0].cpu().numpy(), rtol=2e-6)
    assert torch.equal(a[2], c[2])
    ok = (np.abs(oc.soft_argmax(torch.from_numpy(batch["heatmaps"]).half().float())[0].numpy() % 1 - 0.5) > 1e-3).all(-1)
    assert np.abs(a[1].cpu().numpy() - c[1].cpu().numpy())[ok].max() <= 1e-4
    for i in (3, 4, 5):
        # two kernels, two orders of float32 summation: an element may differ by one ulp of half, plus float32 round-off
        # on the scale of its tile's largest gradient where terms cancel (the bound the float32 kernels meet among themselves)
        d = (a[i].float() - c[i].float()).abs()
        ulp = torch.maximum(c[i].float().abs(), torch.tensor(6.1e-5, device=d.device)) * 2.0 ** -10
        tile_max = c[i].float().abs().amax(dim=(-2, -1), keepdim=True)
        assert bool((d <= ulp + 1e-5 * tile_max).all()) and (a[i] == c[i]).float().mean().item() > 0.99, i
    # forward only (no gradients): the same losses and decode
    e = ops.fusion_loss_f16(half["heatmaps"], half["offsets"], half["variances"], None, vis, kps, None, scale,
                            float(cfg.input_size[0]), float(cfg.input_size[1]), loss_fn.lambdas, float(cfg.sigma), float(cfg.sigma), True, sk,
                            False, True, alpha, fw, 2, DF)
    assert torch.equal(e[0], a[0]) and torch.equal(e[1], a[1]) and torch.equal(e[2], a[2])


def test_softplus_mean_matches_torch(gb):
    """The head-side half of the per-tile-mean path: mean_N(softplus(raw)) and its backward against torch
    (nn.Softplus beta 1 threshold 20, then .mean(dim=(2, 3)) — fusion_head.py:245-251, :467-478)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    raw = (torch.randn(5, 17, 64, 48, generator=g, device="cuda") * 4.0)
    raw[0, 0, 0, :4] = torch.tensor([25.0, 20.0, -30.0, 19.999], device="cuda")        # both sides of the threshold
    a = raw.clone().requires_grad_(True)
    m = gb.softplus_mean(a)
    up = torch.randn(5, 17, generator=g, device="cuda")
    (m * up).sum().backward()
    b = raw.clone().double().requires_grad_(True)
    mt = torch.nn.functional.softplus(b, beta=1.0, threshold=20.0).mean(dim=(2, 3))
    (mt * up.double()).sum().backward()
    np.testing.assert_allclose(m.detach().cpu().numpy(), mt.detach().cpu().numpy(), rtol=2e-6)
    gw = b.grad.cpu().numpy()
    assert np.abs(a.grad.cpu().numpy() - gw).max() <= 2e-6 * np.abs(gw).max()


# ---------------------------------------------------------------------- variance branch handed over as per-tile means
@pytest.mark.parametrize("name", NAMES)
def test_step_with_variance_means(gb, name):
    """a head that reduces V to mean_N(V) in its last convolution's epilogue.  The loss uses V only through
    that mean, so the step with the means must give the step with the maps: same losses, same d/d(heatmaps, offsets),
    and d/d(mean) = N * d/dV_i."""
    cfg, batch, g = goldens.load(name)
    pairs = [v for p in oc.skeleton_for(cfg.K) for v in p]
    a, f = torch.tensor(0.5).cuda(), torch.tensor(0.6224593312018546).cuda()
    common = (float(cfg.input_size[0]), float(cfg.input_size[1]), list(oc.DEFAULT_LAMBDAS), cfg.sigma, cfg.sigma, True, pairs, True, True, a, f, 2, 3)
    hm, off, var = dev(batch["heatmaps"]), dev(batch["offsets"]), dev(batch["variances"])
    # (both calls take the kernel of their shape: the persistent step kernel for 64x48, the one-CTA-per-tile kernel otherwise)
    full = gb.fusion_loss(hm, off, var, None, dev(batch["vis"]), dev(batch["kps"]), None, None, *common)
    vm = var.double().mean(dim=(2, 3)).float()
    res = gb.fusion_step_vmean(hm, off, vm, None, dev(batch["vis"]), dev(batch["kps"]), None, None, *common)
    np.testing.assert_allclose(res[0].cpu().numpy(), full[0].cpu().numpy(), rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(res[0].cpu().numpy(), g["loss_f32"], rtol=LOSS_RTOL, atol=1e-9)
    assert torch.equal(res[1], full[1]) and torch.equal(res[2], full[2])           # the heatmap / offset gradients do not see V
    assert torch.equal(res[4], full[4]) and torch.equal(res[5], full[5])
    n = cfg.H * cfg.W
    np.testing.assert_allclose(res[3].cpu().numpy(), full[3][:, :, 0, 0].cpu().numpy() * n, rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(res[3].cpu().numpy(), g["grad_var_tile"] * n, rtol=1e-4, atol=1e-12)
