"""CPU: the per-tile closed-form loss/gradient algebra that the CUDA kernel
implements equals the oracle's autograd (float64, so only algebra errors show)."""
import numpy as np
import pytest
import torch

from oracle import heatmap_codec as oc
from tests import closed_form, goldens, synth


@pytest.mark.parametrize("name", list(synth.CONFIGS))
@pytest.mark.parametrize("utw", [True, False])
def test_closed_form_matches_autograd_f64(name, utw):
    cfg, batch, g = goldens.load(name)
    d = lambda k: torch.from_numpy(batch[k]).double()
    args = (d("heatmaps"), d("offsets"), d("variances"), d("target"), d("weight"), d("kps"))
    losses, grads = oc.fusion_loss_and_grads(*args, input_size=cfg.input_size, target_sigma=cfg.sigma,
                                             use_target_weight=utw)
    L, gh, go, gv = closed_form.loss_closed_form(*args, cfg.input_size, sigma=cfg.sigma, use_target_weight=utw)
    ref = np.array([float(losses[k]) for k in oc.LOSS_KEYS])
    np.testing.assert_allclose(L.numpy(), ref, rtol=1e-12)
    for got, want in ((gh, grads["heatmaps"]), (go, grads["offsets"]), (gv, grads["variances"])):
        scale = want.abs().max()
        assert (got - want).abs().max() <= 1e-12 * scale
