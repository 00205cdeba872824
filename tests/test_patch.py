"""patch_reference(): the rebind of the reference's entry points.

CPU part: rebinding / restoring on stand-in modules that carry the reference's class and
attribute names (the real reference is used instead when /root/reference is importable, i.e. in
the build container), and the numpy weight rule of the DataLoader-side placeholder against the
oracle encoder.  GPU part: the rebound methods give what the package's own modules give."""
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import heatmap_codec as oc
from tests import synth


def _stand_in_modules():
    fh = types.ModuleType("models.fusion_head")
    pe = types.ModuleType("models.pose_estimator")
    cd = types.ModuleType("datasets.coco_dataset")

    class FusionPoseLoss(nn.Module):                       # attribute names of fusion_head.py:608-635
        def __init__(self, target_sigma=2.0):
            super().__init__()
            self.heatmap_weight, self.offset_weight, self.peak_weight = 1.0, 1.0, 0.5
            self.variance_weight, self.overlap_weight, self.shape_weight = 0.1, 0.05, 0.05
            self.use_target_weight = True
            self.gaussian_constraint = types.SimpleNamespace(target_sigma=target_sigma)

        def forward(self, *a, **k):
            return "original-loss"

    class _Local(nn.Module):
        local_radius = 2

    class _Refine(nn.Module):
        def __init__(self):
            super().__init__()
            self.alpha = nn.Parameter(torch.tensor(0.5))
            self.local_refine = _Local()

    class HeatmapRegressionHead(nn.Module):
        def __init__(self):
            super().__init__()
            self.use_subpixel_refinement = True
            self.subpixel_refine = _Refine()

        def decode(self, outputs, apply_offset=True):
            return "original-decode"

    class PoseEstimator(nn.Module):
        def inference(self, x, flip=True, flip_pairs=None):
            return "original-inference"

        @staticmethod
        def decode_heatmaps(heatmaps, shift=True):
            return "original-argmax"

    class COCOPoseDataset:
        def _generate_target(self, keypoints, keypoints_visible):
            return "original-encode"

    fh.FusionPoseLoss, fh.HeatmapRegressionHead = FusionPoseLoss, HeatmapRegressionHead
    pe.PoseEstimator = PoseEstimator
    cd.COCOPoseDataset = COCOPoseDataset
    return fh, pe, cd


def test_patch_and_unpatch_rebind_the_five_entry_points():
    from infantposeestimation_gaussianbias_b200 import patch
    fh, pe, cd = _stand_in_modules()
    saved = patch.patch_reference(fh, pe, cd, encode_on_device=True)
    assert set(saved) == {"FusionPoseLoss.forward", "HeatmapRegressionHead.decode", "PoseEstimator.inference",
                          "PoseEstimator.decode_heatmaps", "COCOPoseDataset._generate_target"}
    assert fh.FusionPoseLoss.forward is patch._loss_forward
    assert cd.COCOPoseDataset._generate_target is patch._encode_placeholder
    assert isinstance(pe.PoseEstimator.__dict__["decode_heatmaps"], staticmethod)
    patch.unpatch_reference()
    assert fh.FusionPoseLoss().forward() == "original-loss"
    assert fh.HeatmapRegressionHead().decode({}) == "original-decode"
    assert pe.PoseEstimator().inference(None) == "original-inference"
    assert pe.PoseEstimator.decode_heatmaps(None) == "original-argmax"
    assert cd.COCOPoseDataset()._generate_target(None, None) == "original-encode"


@pytest.mark.parametrize("name", list(synth.CONFIGS))
def test_placeholder_weights_equal_the_encoder_weights(name):
    from infantposeestimation_gaussianbias_b200 import patch
    cfg = synth.CONFIGS[name]
    batch = synth.make_batch(cfg, seed=11, B=16)
    ds = types.SimpleNamespace(num_keypoints=cfg.K, heatmap_size=cfg.heatmap_size, input_size=cfg.input_size, sigma=cfg.sigma)
    for b in range(16):
        tile, w = patch._encode_placeholder(ds, batch["kps"][b], batch["vis"][b])
        assert tile.shape == (cfg.K, 0, 0) and tile.dtype == np.float32
        assert np.array_equal(w, batch["weight"][b].reshape(cfg.K, 1))
    # default_collate stacks the placeholders into an empty (B,K,0,0) tensor: the loss reads that as "encode on device"
    stacked = torch.stack([torch.from_numpy(patch._encode_placeholder(ds, batch["kps"][b], batch["vis"][b])[0]) for b in range(2)])
    assert stacked.numel() == 0 and stacked.shape[:2] == (2, cfg.K)


@pytest.mark.gpu
def test_patched_methods_run_the_cuda_ops():
    import infantposeestimation_gaussianbias_b200 as pkg
    pkg.load()
    from infantposeestimation_gaussianbias_b200 import patch
    fh, pe, cd = _stand_in_modules()
    patch.patch_reference(fh, pe, cd, encode_on_device=True)
    try:
        cfg = synth.CONFIGS["w32_256x192"]
        batch = synth.make_batch(cfg, seed=2, B=4)
        dev = lambda k: torch.from_numpy(batch[k]).cuda()
        T = lambda k: torch.from_numpy(batch[k])
        outputs = {"heatmaps": dev("heatmaps").requires_grad_(True), "offsets": dev("offsets").requires_grad_(True),
                   "variances": dev("variances").requires_grad_(True), "fusion_weight": torch.sigmoid(torch.tensor(0.5)).cuda()}
        # the DataLoader-side placeholder: empty target, exact weights -> tiles are built in the kernel
        empty = torch.empty(4, cfg.K, 0, 0).cuda()
        loss = fh.FusionPoseLoss(target_sigma=cfg.sigma)
        out = loss(outputs, empty, dev("weight"), dev("kps"), input_size=cfg.input_size, heatmap_size=cfg.heatmap_size)
        out["total_loss"].backward()
        want_l, want_g = oc.fusion_loss_and_grads(T("heatmaps"), T("offsets"), T("variances"), T("target"), T("weight"), T("kps"),
                                                  input_size=cfg.input_size, target_sigma=cfg.sigma)
        for k in oc.LOSS_KEYS:
            np.testing.assert_allclose(float(out[k]), float(want_l[k]), rtol=1e-5, atol=1e-9)
        gh, wh = outputs["heatmaps"].grad.cpu().numpy(), want_g["heatmaps"].numpy()
        assert np.abs(gh - wh).max() <= 1e-5 * np.abs(wh).max()
        # decode through the rebound head method, arg-max through the rebound static method
        head = fh.HeatmapRegressionHead().cuda()
        with torch.no_grad():
            coords, scores = head.decode(outputs)
            kp, mv = pe.PoseEstimator.decode_heatmaps(outputs["heatmaps"].detach())
        wc, ws = oc.fusion_decode(T("heatmaps"), T("offsets"), 0.5, float(torch.sigmoid(torch.tensor(0.5))))
        ok = (np.abs(oc.soft_argmax(T("heatmaps"))[0].numpy() % 1 - 0.5) > 1e-3).all(-1)
        assert np.abs(coords.cpu().numpy() - wc.numpy())[ok].max() <= 1e-4
        assert np.array_equal(scores.cpu().numpy(), ws.numpy())
        wkp, wmv, _ = oc.decode_heatmaps(T("heatmaps"))
        assert np.array_equal(kp.cpu().numpy(), wkp.numpy()) and np.array_equal(mv.cpu().numpy(), wmv.numpy())
    finally:
        patch.unpatch_reference()


def test_patch_binds_on_the_real_reference_when_present():
    """Build container only: the unmodified reference's classes accept the rebind (attribute names exist)."""
    import os
    ref = os.environ.get("GBCODEC_REF", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "models")):
        pytest.skip("reference tree not present (GPU box)")
    from infantposeestimation_gaussianbias_b200 import patch
    sys.path.insert(0, ref)
    try:
        import importlib
        fh = importlib.import_module("models.fusion_head")
        pe = importlib.import_module("models.pose_estimator")
        originals = (fh.FusionPoseLoss.forward, fh.HeatmapRegressionHead.decode, pe.PoseEstimator.inference)
        patch.patch_reference(fh, pe)
        assert fh.FusionPoseLoss.forward is patch._loss_forward
        loss = fh.FusionPoseLoss()            # the reference's constructor; the shim reads these attributes
        for attr in ("heatmap_weight", "offset_weight", "peak_weight", "variance_weight", "overlap_weight", "shape_weight"):
            assert hasattr(loss, attr)
        head = fh.HeatmapRegressionHead(32)
        assert hasattr(head.subpixel_refine, "alpha") and hasattr(head.subpixel_refine.local_refine, "local_radius")
        patch.unpatch_reference()
        assert (fh.FusionPoseLoss.forward, fh.HeatmapRegressionHead.decode, pe.PoseEstimator.inference) == originals
    finally:
        sys.path.remove(ref)
        for m in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
            del sys.modules[m]


# ------------------------------------------------------------------------------------------------
# second generation: utils/postprocess.py, models/losses.py, KeypointMSELoss
# ------------------------------------------------------------------------------------------------
def _genb_stand_ins():
    pp = types.ModuleType("utils.postprocess")
    ls = types.ModuleType("models.losses")
    pe = types.ModuleType("models.pose_estimator")
    from infantposeestimation_gaussianbias_b200 import patch
    for fn in patch._GENB_FUNCTIONS:
        setattr(pp, fn, (lambda name: (lambda *a, **k: f"original-{name}"))(fn))

    class FusedPoseLoss(nn.Module):                          # attribute names of models/losses.py
        def __init__(self, use_target_weight=True, loss_type="mse"):
            super().__init__()
            self.use_target_weight, self.loss_type = use_target_weight, loss_type

        def forward(self, *a, **k):
            return "original"

    class MorphologyShapeLoss(nn.Module):
        def __init__(self, lambda_variance=1.0, lambda_mean=0.5):
            super().__init__()
            self.lambda_variance, self.lambda_mean = lambda_variance, lambda_mean

        def forward(self, *a, **k):
            return "original"

    class OffsetRegressionLoss(nn.Module):
        def __init__(self, loss_type="smoothl1"):
            super().__init__()
            self.criterion = {"smoothl1": nn.SmoothL1Loss, "l1": nn.L1Loss, "mse": nn.MSELoss}[loss_type](reduction="none")

        def forward(self, *a, **k):
            return "original"

    class JointsMSELoss(nn.Module):
        def __init__(self, use_target_weight=True):
            super().__init__()
            self.use_target_weight = use_target_weight

        def forward(self, *a, **k):
            return "original"

    class CombinedLoss(nn.Module):
        def __init__(self, config):
            super().__init__()
            self.heatmap_loss = FusedPoseLoss(True, "mse")
            self.morph_loss = MorphologyShapeLoss(config.LOSS.MORPH_LAMBDA, 0.5)
            self.regression_loss = OffsetRegressionLoss("smoothl1")
            self.w_heatmap, self.w_morph, self.w_reg = 1.0, config.LOSS.MORPH_WEIGHT, config.LOSS.REG_WEIGHT

        def forward(self, *a, **k):
            return "original"

    class KeypointMSELoss(nn.Module):
        def __init__(self, use_target_weight=True):
            super().__init__()
            self.use_target_weight = use_target_weight

        def forward(self, *a, **k):
            return "original"

    ls.FusedPoseLoss, ls.MorphologyShapeLoss, ls.OffsetRegressionLoss = FusedPoseLoss, MorphologyShapeLoss, OffsetRegressionLoss
    ls.JointsMSELoss, ls.CombinedLoss = JointsMSELoss, CombinedLoss
    pe.KeypointMSELoss = KeypointMSELoss
    return pp, ls, pe


def test_genb_patch_and_unpatch():
    from infantposeestimation_gaussianbias_b200 import patch, postprocess
    pp, ls, pe = _genb_stand_ins()
    saved = patch.patch_reference_genb(pp, ls, pe)
    assert len(saved) == 7 + 5 + 1
    assert pp.postprocess_predictions is postprocess.postprocess_predictions and pp.fused_decode is postprocess.fused_decode
    assert ls.FusedPoseLoss.forward.__name__ == "fused" and pe.KeypointMSELoss.forward.__name__ == "keypoint_mse"
    patch.unpatch_reference_genb()
    assert pp.fused_decode() == "original-fused_decode"
    assert ls.CombinedLoss.forward(None) == "original" and pe.KeypointMSELoss.forward(None) == "original"


def test_genb_patch_binds_on_the_real_reference_when_present():
    import os
    ref = os.environ.get("GBCODEC_REF", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "utils")):
        pytest.skip("reference tree not present (GPU box)")
    from infantposeestimation_gaussianbias_b200 import patch
    sys.path.insert(0, ref)
    try:
        import importlib
        pp = importlib.import_module("utils.postprocess")
        ls = importlib.import_module("models.losses")
        pe = importlib.import_module("models.pose_estimator")
        before = (pp.fused_decode, ls.CombinedLoss.forward, pe.KeypointMSELoss.forward)
        patch.patch_reference_genb(pp, ls, pe)
        # the attributes the rebound forwards read exist on the reference's own modules
        cfg = types.SimpleNamespace(LOSS=types.SimpleNamespace(MORPH_LAMBDA=1.2, MORPH_WEIGHT=0.15, REG_WEIGHT=0.6))
        c = ls.CombinedLoss(cfg)
        assert (c.w_heatmap, c.w_morph, c.w_reg) == (1.0, 0.15, 0.6) and c.morph_loss.lambda_variance == 1.2
        assert type(ls.OffsetRegressionLoss("l1").criterion).__name__ == "L1Loss"
        assert ls.FusedPoseLoss().loss_type == "mse" and pe.KeypointMSELoss().use_target_weight is True
        patch.unpatch_reference_genb()
        assert (pp.fused_decode, ls.CombinedLoss.forward, pe.KeypointMSELoss.forward) == before
    finally:
        sys.path.remove(ref)
        for m in [m for m in sys.modules if m in ("models", "utils") or m.startswith("models.") or m.startswith("utils.")]:
            del sys.modules[m]


@pytest.mark.gpu
def test_genb_patched_entry_points_run_the_cuda_ops():
    import infantposeestimation_gaussianbias_b200 as pkg
    pkg.load()
    from infantposeestimation_gaussianbias_b200 import patch
    from oracle import genb
    pp, ls, pe = _genb_stand_ins()
    patch.patch_reference_genb(pp, ls, pe)
    try:
        cfg = synth.CONFIGS["w32_256x192"]
        batch = synth.make_batch(cfg, seed=3, B=4)
        ex = synth.make_genb_extras(cfg, batch, seed=3)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        config = types.SimpleNamespace(LOSS=types.SimpleNamespace(MORPH_LAMBDA=1.2, MORPH_WEIGHT=0.15, REG_WEIGHT=0.6))
        total, parts = ls.CombinedLoss(config)({"heatmaps": dev(ex["pred"]), "coords": dev(ex["coords"])},
                                               {"heatmaps": dev(batch["target"]), "coords": dev(ex["target_coords"]), "weights": dev(batch["weight"])})
        tw, pw = genb.combined_loss({"heatmaps": T(ex["pred"]), "coords": T(ex["coords"])},
                                    {"heatmaps": T(batch["target"]), "coords": T(ex["target_coords"]), "weights": T(batch["weight"])}, 1.2, 0.15, 0.6)
        for k in pw:
            np.testing.assert_allclose(float(parts[k]), float(pw[k]), rtol=1e-5)
        np.testing.assert_allclose(float(ls.OffsetRegressionLoss("l1")(dev(ex["coords"]), dev(ex["target_coords"]), dev(batch["weight"]))),
                                   float(genb.offset_regression_loss(T(ex["coords"]), T(ex["target_coords"]), T(batch["weight"]), "l1")), rtol=1e-5)
        np.testing.assert_allclose(float(pe.KeypointMSELoss()(dev(ex["pred"]), dev(batch["target"]), dev(batch["weight"]))),
                                   float(genb.keypoint_mse_loss(T(ex["pred"]), T(batch["target"]), T(batch["weight"]))), rtol=1e-5)
        c, v = pp.get_max_preds_with_subpixel(dev(batch["heatmaps"]))
        wc, wv = genb.get_max_preds_with_subpixel(T(batch["heatmaps"]))
        assert np.array_equal(c.cpu().numpy(), wc.numpy()) and np.array_equal(v.cpu().numpy(), wv.numpy())
    finally:
        patch.unpatch_reference_genb()
