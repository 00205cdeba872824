"""GPU: BASELINE.json's full sizes (configs[1]: B=1024 64x48 fused step; configs[2]: B=4096 96x72 decode with
flip test + offsets; configs[3]: preemie 128x128 K=13 B=1024) through size-independent properties, since the
CPU oracle needs minutes there:

  * batch-permutation equivariance: shuffling the images permutes coords / gradients and leaves the losses
    unchanged (up to the order of the second-stage sum);
  * a strided sample of the big batch equals what the oracle gives on that sample where the op is per-tile
    (decode, arg-max), and a 16-image slice run alone with the big batch's normalisers equals the big run there;
  * mirror symmetry of the decode: decoding W-flipped maps mirrors the x coordinate;
  * the flip-test decode of (h, flip(h) with channels swapped) equals the plain decode of h;
  * encode -> arg-max round trip at full size: the integer peak of every active tile is ul + centre.
"""
import numpy as np
import pytest
import torch

from oracle import heatmap_codec as oc
from tests import synth

pytestmark = pytest.mark.gpu
ALPHA, FW = 0.5, 0.6224593312018546


@pytest.fixture(scope="module")
def gb():
    import infantposeestimation_gaussianbias_b200 as pkg
    pkg.load()
    from infantposeestimation_gaussianbias_b200 import ops
    return ops


def device_batch(cfg, B, seed, gb, flip=False):
    """Inputs of the bench's shape family, generated on the device (the numpy generator of tests/synth.py needs
    ~1 minute per GiB)."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(seed)
    K, H, W = cfg.K, cfg.H, cfg.W
    in_w, in_h = cfg.input_size
    u = torch.rand(B, K, generator=g, device=dev)
    vis = torch.where(u < 0.15, 0.0, torch.where(u < 0.40, 1.0, 2.0))
    kps = torch.stack(((torch.rand(B, K, generator=g, device=dev) * 1.2 - 0.1) * in_w,
                       (torch.rand(B, K, generator=g, device=dev) * 1.2 - 0.1) * in_h), -1).contiguous()
    jit = kps + torch.randn(B, K, 2, generator=g, device=dev) * (1.5 * in_w / W)
    t, _ = gb.encode(jit, torch.full_like(vis, 2.0), H, W, float(in_w), float(in_h), cfg.sigma)
    hm = t.mul_(torch.rand(B, K, 1, 1, generator=g, device=dev) * 0.9 + 0.3)
    hm.add_(torch.randn(B, K, H, W, generator=g, device=dev), alpha=0.05)
    d = dict(kps=kps, vis=vis, hm=hm)
    d["off"] = torch.randn(B, K, 2, H, W, generator=g, device=dev).mul_(0.3)
    if flip:
        perm = torch.from_numpy(synth.flip_perm(K)).to(dev)
        d["perm"] = perm
        d["hmf"] = (torch.flip(hm[:, perm.long()], dims=[-1]) + 0.02 * torch.randn(B, K, H, W, generator=g, device=dev)).contiguous()
    return d


def step(gb, cfg, d, sl=slice(None), denoms=None, var=None):
    pairs = [v for p in oc.skeleton_for(cfg.K) for v in p]
    a, f = torch.tensor([ALPHA]).cuda(), torch.tensor([FW]).cuda()
    return gb.fusion_loss(d["hm"][sl], d["off"][sl], None if var is None else var[sl], None, d["vis"][sl], d["kps"][sl], denoms, None,
                          float(cfg.input_size[0]), float(cfg.input_size[1]), list(oc.DEFAULT_LAMBDAS), cfg.sigma, cfg.sigma,
                          True, pairs, True, True, a, f, 2, 3)


@pytest.mark.parametrize("name,B", [("w32_256x192", 512), ("hrformer_384x288", 512), ("preemie_256", 512)])
def test_fused_step_is_bit_reproducible(gb, name, B):
    """No float atomics, fixed-order sums: the same inputs give the same bits, launch after launch — losses, every
    gradient, coordinates and scores (CTAs of 6, 9 and 16 warps; the reductions exchange per-warp partials)."""
    cfg = synth.CONFIGS[name]
    d = device_batch(cfg, B, 23, gb)
    var = torch.nn.functional.softplus(torch.randn(B, cfg.K, cfg.H, cfg.W, device="cuda"))
    first = [t.clone() for t in step(gb, cfg, d, var=var)[:6]]
    for _ in range(6):
        again = step(gb, cfg, d, var=var)[:6]
        for a, b in zip(first, again):
            assert torch.equal(a, b)


@pytest.mark.parametrize("name,B", [("w32_256x192", 1024), ("hrformer_384x288", 512), ("preemie_256", 1024)])
def test_fused_step_full_size_properties(gb, name, B):
    cfg = synth.CONFIGS[name]
    d = device_batch(cfg, B, 11, gb)
    var = torch.nn.functional.softplus(torch.randn(B, cfg.K, cfg.H, cfg.W, device="cuda"))
    losses, ghm, goff, gvar, coords, scores = step(gb, cfg, d, var=var)[:6]
    assert torch.isfinite(losses).all() and torch.isfinite(ghm).all()
    # (1) permutation of the images
    p = torch.randperm(B, device="cuda")
    dp = {k: v[p].contiguous() for k, v in d.items()}
    l2, ghm2, goff2, gvar2, coords2, scores2 = step(gb, cfg, dp, var=var[p].contiguous())[:6]
    np.testing.assert_allclose(l2.cpu().numpy(), losses.cpu().numpy(), rtol=2e-6)
    assert torch.equal(coords2, coords[p]) and torch.equal(scores2, scores[p])
    assert torch.equal(ghm2, ghm[p]) and torch.equal(gvar2, gvar[p]) and torch.equal(goff2, goff[p])
    # (2) a 16-image slice alone, given the big batch's normalisers, reproduces the big run on that slice
    pairs = [v for q in oc.skeleton_for(cfg.K) for v in q]
    den = gb.loss_denominators(d["vis"], d["kps"], False, cfg.H, cfg.W, float(cfg.input_size[0]), float(cfg.input_size[1]), cfg.sigma, pairs)
    sl = slice(B // 2, B // 2 + 16)
    _, ghm_s, goff_s, gvar_s, coords_s, _ = step(gb, cfg, d, sl, denoms=den, var=var)[:6]
    assert torch.equal(ghm_s, ghm[sl]) and torch.equal(goff_s, goff[sl]) and torch.equal(gvar_s, gvar[sl]) and torch.equal(coords_s, coords[sl])
    # (3) 64 images drawn at random from the batch against the CPU oracle, with the big batch's normalisers (the gathered
    #     images run alone first: same bits as inside the big run, so the oracle comparison speaks for the big run)
    ri = torch.randperm(B, generator=torch.Generator().manual_seed(5))[:64].sort().values.cuda()
    dr = {k: v[ri].contiguous() for k, v in d.items()}
    _, ghm_r, goff_r, gvar_r, coords_r, _ = step(gb, cfg, dr, denoms=den, var=var[ri].contiguous())[:6]
    assert torch.equal(ghm_r, ghm[ri]) and torch.equal(goff_r, goff[ri]) and torch.equal(coords_r, coords[ri])
    T = lambda t: t[ri].cpu()
    target, weight = oc.encode_targets(T(d["kps"]).numpy(), T(d["vis"]).numpy(), cfg.heatmap_size, cfg.input_size, cfg.sigma)
    dn = den.cpu().numpy().astype(np.float64)
    want_l, want_g = oc.fusion_loss_and_grads(T(d["hm"]), T(d["off"]), T(var), torch.from_numpy(target), torch.from_numpy(weight), T(d["kps"]),
                                              input_size=cfg.input_size, target_sigma=cfg.sigma, denominators=(dn[0] + 1e-8, dn[1] + 1e-8))
    wh, gh = want_g["heatmaps"].numpy().astype(np.float64), ghm_r.cpu().numpy().astype(np.float64)
    tile_max = np.abs(wh).max(axis=(2, 3), keepdims=True)
    excess = np.abs(gh - wh) / (1e-5 * np.abs(wh) + 1e-5 * tile_max + 1e-30)      # element-wise, floor = 1e-5 of the tile's largest gradient
    print(f"[parity] {name} B={B}: 64 random images vs oracle, element-wise gradient error {excess.max():.2f}x the bound, "
          f"max-norm {np.abs(gh - wh).max() / np.abs(wh).max():.2e}")
    assert excess.max() <= 1.0
    wc, ws_ = oc.fusion_decode(T(d["hm"]), T(d["off"]), ALPHA, FW)
    okc = (np.abs(oc.soft_argmax(T(d["hm"]))[0].numpy() % 1 - 0.5) > 1e-3).all(-1)
    assert (~okc).sum() <= 0.1 * okc.size and np.abs(coords_r.cpu().numpy() - wc.numpy())[okc].max() <= 1e-4
    # (4) the offset gradient holds at most four taps per channel, the variance gradient is uniform per tile
    assert int((goff != 0).sum(dim=(3, 4)).max()) <= 4
    assert torch.equal(gvar, gvar[:, :, :1, :1].expand_as(gvar))


def test_flip_decode_full_size_properties(gb):
    """configs[2]: HRFormer 96x72, decode with flip test + offset correction, batch 4096."""
    cfg = synth.CONFIGS["hrformer_384x288"]
    B = 4096
    d = device_batch(cfg, B, 12, gb, flip=True)
    a, f = torch.tensor([ALPHA]).cuda(), torch.tensor([FW]).cuda()
    c, s, centre = gb.decode(d["hm"], d["hmf"], d["perm"].int(), d["off"], a, f, 2, 3)
    assert torch.isfinite(c).all()
    # (1) a strided sample against the oracle (flip average -> fusion decode)
    idx = torch.arange(0, B, 257, device="cuda")
    avg = oc.flip_average(d["hm"][idx].cpu(), d["hmf"][idx].cpu())          # COCO flip pairs = synth.flip_perm
    want_c, want_s = oc.fusion_decode(avg, d["off"][idx].cpu(), ALPHA, FW)
    frac = np.abs(oc.soft_argmax(avg)[0].numpy() % 1 - 0.5)
    ok = (frac > 1e-3).all(-1)
    assert np.abs(c[idx].cpu().numpy() - want_c.numpy())[ok].max() <= 1e-4
    np.testing.assert_allclose(s[idx].cpu().numpy(), want_s.numpy(), rtol=0, atol=0)
    # (2) flip-test decode of (h, exact mirror of h with channels swapped) == plain decode of h
    exact = torch.flip(d["hm"][:, d["perm"].long()], dims=[-1]).contiguous()
    c1, s1, _ = gb.decode(d["hm"], exact, d["perm"].int(), d["off"], a, f, 2, 3)
    c0, s0, _ = gb.decode(d["hm"], None, None, d["off"], a, f, 2, 3)
    assert torch.equal(c1, c0) and torch.equal(s1, s0)            # (h + h) / 2 == h exactly
    # (3) mirror symmetry without offsets: decoding the W-flipped maps mirrors x (same sums in another order)
    cm, sm, _ = gb.decode(torch.flip(d["hm"], dims=[-1]).contiguous(), None, None, None, a, None, 2, 1)
    cp, sp, _ = gb.decode(d["hm"], None, None, None, a, None, 2, 1)
    assert torch.equal(sm, sp)
    glob, _, _ = gb.decode(d["hm"], None, None, None, None, None, 0, 0)
    away = ((glob - glob.round()).abs() - 0.5).abs().min(dim=-1).values > 1e-3      # rounding-boundary tiles excluded (H1)
    assert ((cm[..., 0] + cp[..., 0] - (cfg.W - 1)).abs()[away].max() <= 1e-4) and ((cm[..., 1] - cp[..., 1]).abs()[away].max() <= 1e-4)


def test_encode_argmax_round_trip_full_size(gb):
    """Every active encoded tile has its single maximum 1.0 at ul + centre; the arg-max finds it (first-max rule
    irrelevant: the peak is unique), at B = 4096."""
    cfg = synth.CONFIGS["w32_256x192"]
    rng = np.random.default_rng(5)
    kps, vis = synth.make_keypoints(cfg, rng, 4096)
    target, weight = gb.encode(torch.from_numpy(kps).cuda(), torch.from_numpy(vis).cuda(), cfg.H, cfg.W, 192.0, 256.0, 2.0)
    c, v, idx = gb.decode_argmax(target, 0)
    mu = kps.astype(np.float64) / 4.0
    ul = np.trunc(mu - 6.0).astype(np.int64)
    peak = ul + 6
    inside = (peak[..., 0] >= 0) & (peak[..., 0] < cfg.W) & (peak[..., 1] >= 0) & (peak[..., 1] < cfg.H) & (weight.cpu().numpy()[..., 0] > 0)
    got = c.cpu().numpy().astype(np.int64)
    assert np.array_equal(got[inside], peak[inside])
    assert np.all(v.cpu().numpy()[inside] == 1.0)
    empty = weight.cpu().numpy()[..., 0] == 0
    assert np.all(v.cpu().numpy()[empty] == 0.0) and np.all(idx.cpu().numpy()[empty] == 0)
