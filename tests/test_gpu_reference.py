"""The REAL reference (unmodified tree: $GBCODEC_REF, /root/reference or baseline/_ref), stock vs patched, on the GPU.

What `north_star` promises — "models/pose_estimator.py, train.py and validate.py use them unchanged" — run, not
asserted: the reference's own `build_model` HRNet-W32 + fusion head on CUDA, the body of train.py:159-183 (fp32 and
autocast + GradScaler), `PoseEstimator.inference(flip=True)` (pose_estimator.py:275-329), validate.py:64-119's loop
body, and `COCOPoseDataset._generate_target` (coco_dataset.py:185-250), each once with the stock methods and once
after `patch_reference()`.  Every comparison prints what it measured (run with -s to see it; the gpurun log is kept
under profiles/).
"""
import copy
import types

import numpy as np
import pytest
import torch

from tests import synth
from tests.refload import Reference

pytestmark = pytest.mark.gpu

CFG = synth.CONFIGS["w32_256x192"]
B = 4


@pytest.fixture(scope="module")
def world():
    r = Reference()
    if not r.present:
        pytest.skip("reference tree not present (run tools/install_reference.py in the build container)")
    import infantposeestimation_gaussianbias_b200 as pkg
    pkg.load()
    from infantposeestimation_gaussianbias_b200 import patch
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    with r as ref:
        cfg = ref.config.get_config()
        cfg.model.pretrained = False
        torch.manual_seed(0)
        model = ref.models.build_model(cfg).cuda()
        g = torch.Generator().manual_seed(3)
        imgs = torch.randn(B, 3, CFG.input_size[1], CFG.input_size[0], generator=g).cuda()
        # The random-init network in eval mode (BatchNorm on its initial running statistics) puts out heatmaps with a
        # standard deviation of ~140 and offsets of ~130 px; at that scale the reference's own float32 decode is 6e-4 px
        # away from its float64 evaluation (the bilinear offset read amplifies the last bits of the soft-argmax), so a
        # 1e-4 px comparison would measure float32 noise.  The model used for inference / validation gets the last
        # 1x1 convolutions of the heatmap and offset branches scaled so that its outputs sit where a trained head's do
        # (heatmaps of order 1, offsets a fraction of a pixel).  Nothing else is touched.
        model_eval = copy.deepcopy(model).eval()
        with torch.no_grad():
            raw = model_eval(imgs)
            for branch, key, want_std in ((model_eval.head.heatmap_branch, "heatmaps", 1.0), (model_eval.head.offset_branch, "offsets", 0.3)):
                f = want_std / float(raw[key].std())
                branch[3].weight.mul_(f); branch[3].bias.mul_(f)
            cal = model_eval(imgs)
        print(f"[real-reference] eval-mode outputs: heatmaps std {float(raw['heatmaps'].std()):.3g} -> {float(cal['heatmaps'].std()):.3g}, "
              f"offsets std {float(raw['offsets'].std()):.3g} -> {float(cal['offsets'].std()):.3g}")
        batch = synth.make_batch(CFG, seed=5, B=B)
        # targets from the reference's own encoder (coco_dataset.py:185-250), one sample at a time like __getitem__
        ds = object.__new__(ref.coco_dataset.COCOPoseDataset)
        ds.num_keypoints, ds.sigma = CFG.K, CFG.sigma
        ds.heatmap_size, ds.input_size = np.array(CFG.heatmap_size), np.array(CFG.input_size)      # coco_dataset.py:60-61
        enc = [ds._generate_target(batch["kps"][b], batch["vis"][b]) for b in range(B)]
        targets = torch.from_numpy(np.stack([e[0] for e in enc])).cuda()
        weights = torch.from_numpy(np.stack([e[1] for e in enc])).cuda()
        kps = torch.from_numpy(batch["kps"]).cuda()
        yield types.SimpleNamespace(ref=ref, cfg=cfg, model=model, model_eval=model_eval, imgs=imgs, targets=targets, weights=weights, kps=kps,
                                    batch=batch, patch=patch, ds=ds)
        patch.unpatch_reference()


def _train_step(w, model, fp16, scale=64.0):
    """train.py:159-183 for one batch, without the optimizer step (the gradients are what is compared).
    The backbone's BatchNorm layers run on their running statistics (frozen-BN fine-tuning): with batch statistics over
    4 random images the random-init HRNet is chaotic — on the CPU two copies of the stock model, same input, differ by
    O(1) in the features, and a 1e-6 input perturbation does the same — so stock and patched steps would not see the
    same heatmaps.  The head (where the codec's inputs come from) stays in train mode."""
    model.train()
    model.backbone.eval()
    model.zero_grad(set_to_none=True)
    scaler = torch.amp.GradScaler("cuda", enabled=fp16, init_scale=scale)
    with torch.autocast("cuda", enabled=fp16):
        outputs = model(w.imgs, w.targets, w.weights, gt_keypoints=w.kps, input_size=w.cfg.data.input_size)
        loss = outputs["loss"]
    scaler.scale(loss).backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().float().clone() / (scale if fp16 else 1.0) for n, p in model.named_parameters() if p.grad is not None}
    return {k: float(v) for k, v in outputs["losses"].items()}, grads


def _head_last_convs(grads):
    return {n: g for n, g in grads.items() if n.startswith("head.") and (".3.weight" in n or ".3.bias" in n)}


def _report(tag, **kw):
    print(f"[real-reference] {tag}: " + ", ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in kw.items()))


def test_train_step_fp32_stock_vs_patched(world):
    w = world
    model = copy.deepcopy(w.model)
    w.patch.unpatch_reference()
    want_l, want_g = _train_step(w, copy.deepcopy(w.model), fp16=False)
    w.patch.patch_reference(w.ref.fusion_head, w.ref.pose_estimator)
    try:
        got_l, got_g = _train_step(w, copy.deepcopy(w.model), fp16=False)
    finally:
        w.patch.unpatch_reference()
    assert set(got_l) == set(want_l) and len(want_l) == 7
    worst_l = max(abs(got_l[k] - want_l[k]) / max(abs(want_l[k]), 1e-12) for k in want_l)
    for k in want_l:
        np.testing.assert_allclose(got_l[k], want_l[k], rtol=1e-5, atol=1e-9, err_msg=k)
    heads = _head_last_convs(want_g)
    assert len(heads) >= 6, sorted(heads)
    worst_h = 0.0
    for n, g in heads.items():
        err = float((got_g[n] - g).abs().max()) / max(float(g.abs().max()), 1e-30)
        worst_h = max(worst_h, err)
        assert err <= 1e-4, (n, err)
    # every parameter of the network (backbone included: the gradients flowed through the head into stock PyTorch)
    worst_all = max(float((got_g[n] - g).abs().max()) / max(float(g.abs().max()), 1e-30) for n, g in want_g.items())
    assert set(got_g) == set(want_g) and worst_all <= 1e-3
    _report("train step fp32", loss_rel_err=worst_l, head_last_conv_grad_maxnorm_err=worst_h, all_params_grad_maxnorm_err=worst_all,
            n_params=len(want_g), total_loss=want_l["total_loss"])


def test_train_step_encode_on_device_placeholder(world):
    """DataLoader-side placeholder (empty target + exact weights): the tiles are built inside the kernel."""
    w = world
    model = copy.deepcopy(w.model)
    w.patch.unpatch_reference()
    want_l, want_g = _train_step(w, copy.deepcopy(w.model), fp16=False)
    w.patch.patch_reference(w.ref.fusion_head, w.ref.pose_estimator, w.ref.coco_dataset, encode_on_device=True)
    try:
        enc = [w.ds._generate_target(w.batch["kps"][b], w.batch["vis"][b]) for b in range(B)]       # rebound: placeholder
        assert enc[0][0].shape == (CFG.K, 0, 0)
        w2 = copy.copy(w)
        w2.targets = torch.from_numpy(np.stack([e[0] for e in enc])).cuda()
        w2.weights = torch.from_numpy(np.stack([e[1] for e in enc])).cuda()
        assert torch.equal(w2.weights, w.weights)
        got_l, got_g = _train_step(w2, copy.deepcopy(w.model), fp16=False)
    finally:
        w.patch.unpatch_reference()
    for k in want_l:
        np.testing.assert_allclose(got_l[k], want_l[k], rtol=1e-5, atol=1e-9, err_msg=k)
    worst = max(float((got_g[n] - g).abs().max()) / max(float(g.abs().max()), 1e-30) for n, g in _head_last_convs(want_g).items())
    assert worst <= 1e-4
    _report("train step, encode on device", head_last_conv_grad_maxnorm_err=worst)


def test_train_step_variance_means_stock_vs_patched(world):
    """patch_reference(variance_means=True) — the head's variance branch ends in one kernel (Softplus ->
    mean over the tile) instead of the Softplus module, outputs['variances'] is (B,K), the loss takes the map-less step
    (16N instead of 24N bytes per tile) and neither the variance map nor its gradient map exists.  Same losses, same
    gradients — including the variance branch's own convolution weights, which the gradient reaches through
    gbcodec_softplus_mean_backward_f32."""
    w = world
    w.patch.unpatch_reference()
    want_l, want_g = _train_step(w, copy.deepcopy(w.model), fp16=False)
    w.patch.patch_reference(w.ref.fusion_head, w.ref.pose_estimator, variance_means=True)
    try:
        model = copy.deepcopy(w.model)
        with torch.no_grad():
            model.train(); model.backbone.eval()
            assert model(w.imgs)["variances"].shape == (B, CFG.K)
        got_l, got_g = _train_step(w, copy.deepcopy(w.model), fp16=False)
    finally:
        w.patch.unpatch_reference()
    worst_l = max(abs(got_l[k] - want_l[k]) / max(abs(want_l[k]), 1e-12) for k in want_l)
    for k in want_l:
        np.testing.assert_allclose(got_l[k], want_l[k], rtol=1e-5, atol=1e-9, err_msg=k)
    heads = _head_last_convs(want_g)
    worst_h = max(float((got_g[n] - g).abs().max()) / max(float(g.abs().max()), 1e-30) for n, g in heads.items())
    var_w = [n for n in heads if "variance_branch" in n]
    assert var_w and worst_h <= 1e-4
    worst_all = max(float((got_g[n] - g).abs().max()) / max(float(g.abs().max()), 1e-30) for n, g in want_g.items())
    assert worst_all <= 1e-3
    _report("train step, variance branch as per-tile means", loss_rel_err=worst_l, head_last_conv_grad_maxnorm_err=worst_h,
            all_params_grad_maxnorm_err=worst_all)


def test_train_step_autocast_gradscaler_stock_vs_patched(world):
    """train.py:171-183 with cfg.train.fp16 (the reference's default).  The stock loss runs on the float16 maps under
    autocast (elementwise work in half, softmax / mse / smooth_l1 in float); the codec reads the same float16 maps and
    computes in float32 throughout, so the two differ by the stock path's half-precision rounding, not by 1e-5: the
    bound here is 2e-3 on the losses and 2e-2 max-norm on the head's last-conv gradients, and the test prints what it
    saw.  The float32 test above is the parity gate."""
    w = world
    model = copy.deepcopy(w.model)
    w.patch.unpatch_reference()
    want_l, want_g = _train_step(w, copy.deepcopy(w.model), fp16=True)
    w.patch.patch_reference(w.ref.fusion_head, w.ref.pose_estimator)
    try:
        got_l, got_g = _train_step(w, copy.deepcopy(w.model), fp16=True)
    finally:
        w.patch.unpatch_reference()
    worst_l = max(abs(got_l[k] - want_l[k]) / max(abs(want_l[k]), 1e-12) for k in want_l)
    heads = _head_last_convs(want_g)
    bad_stock = sum(int((~torch.isfinite(g)).sum()) for g in want_g.values())
    bad_ours = sum(int((~torch.isfinite(g)).sum()) for g in got_g.values())
    worst_h = max(float((got_g[n] - g).abs().max()) / max(float(g.abs().max()), 1e-30) for n, g in heads.items())
    _report("train step autocast+GradScaler", loss_rel_err=worst_l, head_last_conv_grad_maxnorm_err=worst_h,
            nonfinite_grad_entries_stock=bad_stock, nonfinite_grad_entries_patched=bad_ours)
    assert all(np.isfinite(v) for v in got_l.values())
    assert bad_ours <= bad_stock, "the patched step overflowed where the stock step did not"
    if bad_stock:
        pytest.skip(f"the stock autocast step itself produced {bad_stock} non-finite gradient entries at this loss scale "
                    f"(GradScaler would skip the step); losses agree to {worst_l:.2g}")
    assert worst_l <= 2e-3 and worst_h <= 2e-2


def _h1_ok(heatmaps):
    """tiles whose soft-argmax is not within 1e-3 px of a half-integer (the window centre is a rounding)."""
    from oracle import heatmap_codec as oc
    c = oc.soft_argmax(heatmaps.float().cpu())[0].numpy()
    return (np.abs(c % 1 - 0.5) > 1e-3).all(-1)


def test_inference_flip_stock_vs_patched(world):
    w = world
    model = copy.deepcopy(w.model_eval).eval()
    flip_pairs = w.cfg.data.flip_pairs
    w.patch.unpatch_reference()
    with torch.no_grad():
        want_c, want_s = model.inference(w.imgs, flip=True, flip_pairs=flip_pairs)
        want_c0, want_s0 = model.inference(w.imgs, flip=False)
        # the averaged heatmap, for the H1 mask only
        h = model(w.imgs)["heatmaps"]
        hf = torch.flip(model(torch.flip(w.imgs, dims=[-1]))["heatmaps"], dims=[-1])
        hf2 = hf.clone()
        for a, b in flip_pairs:
            hf2[:, a], hf2[:, b] = hf[:, b], hf[:, a]
        ok = _h1_ok((h + hf2) / 2)
        ok0 = _h1_ok(h)
    w.patch.patch_reference(w.ref.fusion_head, w.ref.pose_estimator)
    try:
        with torch.no_grad():
            got_c, got_s = model.inference(w.imgs, flip=True, flip_pairs=flip_pairs)
            got_c0, got_s0 = model.inference(w.imgs, flip=False)
    finally:
        w.patch.unpatch_reference()
    for tag, gc, gs, wc, ws, m in (("flip", got_c, got_s, want_c, want_s, ok), ("no flip", got_c0, got_s0, want_c0, want_s0, ok0)):
        err = (gc - wc).abs().max(-1).values.cpu().numpy()
        excluded = int((~m).sum())
        _report(f"inference {tag}", tiles=m.size, h1_excluded=excluded, max_err_px=float(err[m].max()), max_err_px_all=float(err.max()))
        assert excluded <= 0.1 * m.size
        assert err[m].max() <= 1e-4
        assert torch.equal(gs, ws)


def test_validate_loop_body_stock_vs_patched(world):
    """validate.py:64-119 for one batch: flip-test inference, the logged loss under no_grad, heatmap px -> input px ->
    original image with utils.transforms.transform_preds."""
    w = world
    model = copy.deepcopy(w.model_eval).eval()
    cfg = w.cfg
    transform_preds = w.ref.module("validate").transform_preds                  # validate.py:31-36
    centers = np.stack([np.array([320.0 + 7 * i, 240.0 - 5 * i], dtype=np.float32) for i in range(B)])
    scales = np.stack([np.array([150.0 + 10 * i, 200.0 + 10 * i], dtype=np.float32) for i in range(B)])     # crop size in image px

    def body():
        with torch.no_grad():
            pk, ps = model.inference(w.imgs, flip=True, flip_pairs=cfg.data.flip_pairs)
            outputs = model(w.imgs, w.targets, w.weights, gt_keypoints=w.kps, input_size=cfg.data.input_size)
            loss = outputs["loss"].item()
        pk, ps = pk.cpu().numpy(), ps.cpu().numpy()
        pk[:, :, 0] *= cfg.data.input_size[0] / cfg.data.heatmap_size[0]
        pk[:, :, 1] *= cfg.data.input_size[1] / cfg.data.heatmap_size[1]
        for i in range(B):
            for k in range(cfg.data.num_keypoints):
                pk[i, k] = transform_preds(pk[i, k], centers[i], scales[i], cfg.data.input_size)
        return pk, ps, loss

    w.patch.unpatch_reference()
    want_k, want_s, want_loss = body()
    with torch.no_grad():
        h = model(w.imgs)["heatmaps"]
        hf = torch.flip(model(torch.flip(w.imgs, dims=[-1]))["heatmaps"], dims=[-1])
        hf2 = hf.clone()
        for a, b in cfg.data.flip_pairs:
            hf2[:, a], hf2[:, b] = hf[:, b], hf[:, a]
        ok = _h1_ok((h + hf2) / 2)
    w.patch.patch_reference(w.ref.fusion_head, w.ref.pose_estimator)
    try:
        got_k, got_s, got_loss = body()
    finally:
        w.patch.unpatch_reference()
    np.testing.assert_allclose(got_loss, want_loss, rtol=1e-5)
    assert np.array_equal(got_s, want_s)
    # 1e-4 heatmap px = 4e-4 input px, times crop size / input size (<= 230 / 192 here) in the original image
    tol = 1e-4 * 4 * 230 / 192 * 1.5
    err = np.abs(got_k - want_k).max(-1)
    _report("validate loop body", loss_rel_err=abs(got_loss - want_loss) / abs(want_loss), max_err_image_px=float(err[ok].max()),
            h1_excluded=int((~ok).sum()), tol_image_px=tol)
    assert err[ok].max() <= tol


def test_generate_target_stock_vs_device_encoder(world):
    """COCOPoseDataset._generate_target (the stock method, numpy) against the batched device encoder."""
    from infantposeestimation_gaussianbias_b200 import generate_heatmaps
    w = world
    got_t, got_w = generate_heatmaps(w.kps, torch.from_numpy(w.batch["vis"]).cuda(), CFG.heatmap_size, CFG.input_size, CFG.sigma)
    assert torch.equal(got_w, w.weights)
    gt, wt = got_t.cpu().numpy(), w.targets.cpu().numpy()
    assert np.array_equal(gt != 0, wt != 0)
    np.testing.assert_allclose(gt, wt, rtol=1e-6, atol=0)
    _report("encode", tiles=B * CFG.K, max_rel_err=float((np.abs(gt - wt) / np.maximum(wt, 1e-30))[wt != 0].max()))


def test_genb_combined_loss_under_autocast_stock_vs_float16_kernels(world):
    """The second generation's CombinedLoss (models/losses.py:205-290) under torch.autocast with float16 heatmaps: the STOCK
    reference module against this package's, which hands the half maps to gbcodec_combined_loss_f16 as they are.  Under
    autocast the reference does its elementwise work in half and its reductions / criteria in float; the kernel computes
    everything in float32 on the up-cast values.  Against the stock autocast step the bound is therefore half-precision
    round-off (on the CPU, where autocast leaves `sum` in half, the stock losses are 1e-3 from their own float32
    evaluation; on CUDA `sum` is promoted and the stock step is much closer); against the stock module's FLOAT32 evaluation
    of the same up-cast maps the bound is the usual one.  The distances are printed."""
    ref = world.ref
    RL = ref.module("models.losses")
    from infantposeestimation_gaussianbias_b200 import losses as L
    K = CFG.K
    cfg_l = types.SimpleNamespace(LOSS=types.SimpleNamespace(MORPH_LAMBDA=1.2, MORPH_WEIGHT=0.15, REG_WEIGHT=0.6))
    g = torch.Generator(device="cuda").manual_seed(11)
    pred = (world.targets * 0.8 + 0.05 * torch.rand(world.targets.shape, generator=g, device="cuda")).half()
    coords = torch.rand(B, K, 2, generator=g, device="cuda") * 40.0
    tcoords = coords + torch.randn(B, K, 2, generator=g, device="cuda")
    tg = {"heatmaps": world.targets, "coords": tcoords, "weights": world.weights}
    scale = 256.0

    def run(module, p, autocast):
        a = p.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            total, parts = module({"heatmaps": a, "coords": coords}, tg)
        (total * scale).backward()
        return {k: float(v.detach()) for k, v in parts.items()}, a.grad

    stock16, g_stock16 = run(RL.CombinedLoss(cfg_l), pred, True)
    stock32, g_stock32 = run(RL.CombinedLoss(cfg_l), pred.float(), False)          # the float32 evaluation of the same maps
    ours16, g_ours16 = run(L.CombinedLoss(cfg_l), pred, True)
    assert g_ours16.dtype == torch.float16 and g_stock16.dtype == torch.float16
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-12)
    worst_vs_stock16 = max(rel(ours16[k], stock16[k]) for k in stock16)
    worst_vs_f32 = max(rel(ours16[k], stock32[k]) for k in stock32)
    stock16_vs_f32 = max(rel(stock16[k], stock32[k]) for k in stock32)
    gmax = float(g_stock32.abs().max())
    d_ours = float((g_ours16.float() - g_stock32).abs().max()) / gmax
    d_stock = float((g_stock16.float() - g_stock32).abs().max()) / gmax
    d_between = float((g_ours16.float() - g_stock16.float()).abs().max()) / gmax
    _report("Gen-B CombinedLoss autocast", loss_rel_ours_vs_stock16=worst_vs_stock16, loss_rel_ours_vs_float32=worst_vs_f32,
            loss_rel_stock16_vs_float32=stock16_vs_f32, grad_maxnorm_ours_vs_float32=d_ours, grad_maxnorm_stock16_vs_float32=d_stock,
            grad_maxnorm_ours_vs_stock16=d_between)
    assert set(ours16) == set(stock16)
    assert worst_vs_f32 <= 1e-5                         # float32 arithmetic on the up-cast maps: the float32 reference's losses
    assert d_ours <= 2.0 ** -10                         # ... and its gradient rounded to half once
    assert worst_vs_stock16 <= 5e-3 and d_between <= 2e-2          # half-precision round-off of the stock autocast step
