"""CPU: the host-side helpers of the ops layer that never touch the device — the memoised call descriptor, the flattened
limb pairs of the loss module, the scalar-argument fast path, the device guard that refuses anything but a CUDA device
(no CPU path), and the float16 routing rule of the Gen-B losses."""
import pytest
import torch

from infantposeestimation_gaussianbias_b200 import _native as N
from infantposeestimation_gaussianbias_b200 import ops
from infantposeestimation_gaussianbias_b200.fusion_head import FusionPoseLoss, SKELETON


def test_descriptor_is_memoised_by_value():
    hm = torch.empty(4, 17, 64, 48)
    loss = FusionPoseLoss(target_sigma=2.0)
    pairs = loss._pairs_flat(17)
    a = ops._desc(hm, 192.0, 256.0, loss.lambdas, 2.0, 2.0, True, pairs)
    b = ops._desc(hm, 192.0, 256.0, list(loss.lambdas), 2.0, 2.0, True, list(pairs))
    assert a is b                                                  # same values -> the same structure, not a new one
    assert (a.B, a.K, a.H, a.W, a.n_pairs) == (4, 17, 64, 48, len(pairs) // 2)
    assert [a.lambdas[i] for i in range(6)] == pytest.approx(loss.lambdas)
    assert [(a.pairs[i][0], a.pairs[i][1]) for i in range(a.n_pairs)] == [p for p in SKELETON if p[0] < 17 and p[1] < 17]
    for other in (ops._desc(torch.empty(5, 17, 64, 48), 192.0, 256.0, loss.lambdas, 2.0, 2.0, True, pairs),
                  ops._desc(hm, 192.0, 256.0, [1, 1, 1, 1, 1, 2.0], 2.0, 2.0, True, pairs),
                  ops._desc(hm, 192.0, 256.0, loss.lambdas, 1.5, 2.0, True, pairs),
                  ops._desc(hm, 192.0, 256.0, loss.lambdas, 2.0, 2.0, False, pairs),
                  ops._desc(hm, 192.0, 256.0, loss.lambdas, 2.0, 2.0, True, pairs[:-2])):
        assert other is not a
    assert ops._desc(hm, 192.0, 256.0, [1, 1, 1, 1, 1, 2.0], 2.0, 2.0, True, pairs).lambdas[5] == 2.0
    # the cache is bounded
    for b_ in range(200):
        ops._desc(torch.empty(b_ + 10, 3, 8, 8), 32.0, 32.0, loss.lambdas, 2.0, 2.0, True, [0, 1])
    assert len(ops._DESC_CACHE) <= 64


def test_limb_pairs_follow_the_channel_count():
    loss = FusionPoseLoss(target_sigma=2.0)
    assert loss._pairs_flat(17) == ops.pairs_flat(loss.pairs_for(17))
    assert loss._pairs_flat(13) == ops.pairs_flat(loss.pairs_for(13))          # fusion_head.py:504-505: pairs beyond K are skipped
    assert loss._pairs_flat(17) == ops.pairs_flat(loss.pairs_for(17))
    assert all(v < 13 for v in loss._pairs_flat(13))
    other = FusionPoseLoss(target_sigma=2.0, skeleton=((0, 1), (1, 2)))
    assert other._pairs_flat(17) == [0, 1, 1, 2]


def test_scalar_arguments():
    like = torch.zeros(3)
    a = torch.tensor([0.5])
    assert ops._scalar("alpha", a, like) is a                                   # nothing to convert: the tensor itself
    b = ops._scalar("alpha", torch.tensor(0.5), like)
    assert b.shape == (1,) and float(b) == 0.5
    p = torch.nn.Parameter(torch.tensor(0.25))
    c = ops._scalar("alpha", p, like)
    assert c.shape == (1,) and not c.requires_grad and float(c) == 0.25         # a learnable scalar is detached
    d = ops._scalar("alpha", torch.tensor([2], dtype=torch.float64), like)
    assert d.dtype == torch.float32 and float(d) == 2.0
    assert ops._scalar("alpha", None, like) is None
    with pytest.raises(RuntimeError):
        ops._scalar("alpha", torch.zeros(2), like)


def test_no_cpu_path():
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops._on_device(torch.device("cpu"))
    with pytest.raises(RuntimeError):
        ops._cuda_f32("heatmaps", torch.zeros(1, 1, 4, 4))
    with pytest.raises(RuntimeError):
        ops._cuda_f16("heatmaps", torch.zeros(1, 1, 4, 4, dtype=torch.float16), (1, 1, 4, 4))


def test_float16_entry_points_are_declared():
    # the float16 calls of both generations are part of the ABI the loader checks (include/gbcodec.h <-> EXPORTS)
    for name in ("gbcodec_fusion_step_f16", "gbcodec_fusion_loss_backward_f16", "gbcodec_combined_loss_f16", "gbcodec_combined_loss_backward_f16"):
        assert name in N.EXPORTS
