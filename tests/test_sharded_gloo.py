"""World-size-2 `gloo` run of the batch-sharded loss on CPU: the host-side logic of
sharded.py (normaliser all-reduce before the local pass, loss all-reduce after it, gradient =
the rank's exact share of the global-batch gradient, shard bounds).  The CUDA ops are replaced
by the oracle through the stand-in hooks ShardedFusionPoseLoss exposes for this purpose."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import heatmap_codec as oc
from tests import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from infantposeestimation_gaussianbias_b200.sharded import ShardedFusionPoseLoss, shard_bounds
        cfg = synth.CONFIGS["w32_256x192"]
        batch = synth.make_batch(cfg, seed=3, B=B)
        lo, hi = shard_bounds(B, rank, world)
        T = lambda k: torch.from_numpy(batch[k][lo:hi].copy())

        def local_denominators(weight, gt, given, H, W, input_size):
            # raw sums, no epsilon: that is what gbcodec_loss_denominators_f32 hands to the all-reduce
            K = gt.shape[1]
            w = weight.reshape(-1, K)
            pairs = oc.skeleton_for(K)
            sp = sum((w[:, i] * w[:, j]).sum() for i, j in pairs)
            return torch.stack([w.sum(), torch.as_tensor(sp, dtype=w.dtype)]).float()

        def local_loss(outputs, target, weight, gt, input_size, den):
            return oc.fusion_loss(outputs["heatmaps"], outputs["offsets"], outputs["variances"], target, weight, gt,
                                  input_size=input_size, denominators=(den[0] + 1e-8, den[1] + 1e-8))

        loss_fn = ShardedFusionPoseLoss(target_sigma=cfg.sigma, local_denominators=local_denominators, local_loss=local_loss)
        outputs = {"heatmaps": T("heatmaps").requires_grad_(True), "offsets": T("offsets").requires_grad_(True),
                   "variances": T("variances").requires_grad_(True)}
        out = loss_fn(outputs, T("target"), T("weight"), T("kps"), input_size=cfg.input_size)
        out["total_loss"].backward()
        q.put((rank, lo, hi, {k: float(out[k]) for k in oc.LOSS_KEYS},
               {k: v.grad.numpy() for k, v in outputs.items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_loss_equals_global_batch():
    B, world = 5, 2                      # uneven split: 3 + 2 images
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=3, B=B)
    T = lambda k: torch.from_numpy(batch[k])
    want_l, want_g = oc.fusion_loss_and_grads(T("heatmaps"), T("offsets"), T("variances"), T("target"), T("weight"), T("kps"),
                                              input_size=cfg.input_size, target_sigma=cfg.sigma)
    covered = 0
    for rank, lo, hi, losses, grads in sorted(got, key=lambda g: g[0]):
        covered += hi - lo
        for k in oc.LOSS_KEYS:      # every rank reports the GLOBAL loss
            assert abs(losses[k] - float(want_l[k])) <= 2e-6 * abs(float(want_l[k])) + 1e-9, (rank, k)
        for k in ("heatmaps", "offsets", "variances"):   # and holds its slice of the global gradient
            w = want_g[k].numpy()[lo:hi]
            assert np.abs(grads[k] - w).max() <= 2e-6 * np.abs(want_g[k].numpy()).max(), (rank, k)
    assert covered == B


def test_shard_bounds_cover_the_batch_once():
    from infantposeestimation_gaussianbias_b200.sharded import shard_bounds
    for B in (1, 2, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


# ------------------------------------------------------------------------------------------------
# second generation: CombinedLoss sharded over the batch
# ------------------------------------------------------------------------------------------------
def _genb_worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from infantposeestimation_gaussianbias_b200.sharded import ShardedCombinedLoss, shard_bounds
        from oracle import genb
        cfg = synth.CONFIGS["w32_256x192"]
        batch = synth.make_batch(cfg, seed=4, B=B)
        ex = synth.make_genb_extras(cfg, batch, seed=4)
        lo, hi = shard_bounds(B, rank, world)

        def local_loss(predictions, targets, norm_batch):
            # the oracle's means run over the local batch: rescale to the global one, as norm_batch does in the kernel
            total, parts = genb.combined_loss(predictions, targets, 1.2, 0.15, 0.6)
            f = predictions["heatmaps"].shape[0] / norm_batch
            return total * f, {k: v * f for k, v in parts.items()}

        loss = ShardedCombinedLoss(None, local_loss=local_loss)
        p = torch.from_numpy(ex["pred"][lo:hi].copy()).requires_grad_(True)
        c = torch.from_numpy(ex["coords"][lo:hi].copy()).requires_grad_(True)
        total, parts = loss({"heatmaps": p, "coords": c},
                            {"heatmaps": torch.from_numpy(batch["target"][lo:hi].copy()), "coords": torch.from_numpy(ex["target_coords"][lo:hi].copy()),
                             "weights": torch.from_numpy(batch["weight"][lo:hi].copy())})
        total.backward()
        q.put((rank, lo, hi, {k: float(v) for k, v in parts.items()}, p.grad.numpy(), c.grad.numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_combined_loss_equals_global_batch():
    from oracle import genb
    B, world = 5, 2                      # uneven shards on purpose
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_genb_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=4, B=B)
    ex = synth.make_genb_extras(cfg, batch, seed=4)
    p = torch.from_numpy(ex["pred"]).requires_grad_(True)
    c = torch.from_numpy(ex["coords"]).requires_grad_(True)
    total, parts = genb.combined_loss({"heatmaps": p, "coords": c}, {"heatmaps": torch.from_numpy(batch["target"]), "coords": torch.from_numpy(ex["target_coords"]),
                                                                      "weights": torch.from_numpy(batch["weight"])}, 1.2, 0.15, 0.6)
    total.backward()
    for rank, lo, hi, losses, gp, gc in got:
        for k, v in parts.items():
            np.testing.assert_allclose(losses[k], float(v), rtol=1e-5)
        np.testing.assert_allclose(gp, p.grad.numpy()[lo:hi], rtol=1e-5, atol=1e-7 * np.abs(p.grad.numpy()).max())
        np.testing.assert_allclose(gc, c.grad.numpy()[lo:hi], rtol=1e-5, atol=1e-9)
