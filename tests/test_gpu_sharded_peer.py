"""GPU: the batch-sharded step whose 2 + 7 scalars travel through NVLink peer-memory mailboxes written
by the kernels (csrc/peer.cu, denoms_kernel / finalize_kernel) instead of NCCL all-reduces.

Two ranks, one process each (gloo carries the 64-byte IPC handles only).  On a box with one GPU both
processes share it — the mailboxes are then mapped through CUDA IPC on the same device and the two
contexts time-slice, which is slow but exercises the same protocol; with >= 2 GPUs each rank has its
own device and the stores cross NVLink.  Checked: losses on every rank == the single-GPU losses of the
whole batch, gradients == that batch's gradients restricted to the shard, several consecutive steps
(sequence numbers / parity slots), no spin time-outs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import heatmap_codec as oc
from tests import synth

pytestmark = pytest.mark.gpu
STEPS = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import infantposeestimation_gaussianbias_b200 as pkg
        pkg.load()
        from infantposeestimation_gaussianbias_b200.sharded import PeerExchange, ShardedFusionPoseLoss, shard_bounds
        cfg = synth.CONFIGS["w32_256x192"]
        peer = PeerExchange(device=dev)
        loss_fn = ShardedFusionPoseLoss(target_sigma=cfg.sigma, peer=peer)
        results = []
        for step in range(STEPS):
            batch = synth.make_batch(cfg, seed=50 + step, B=B)
            lo, hi = shard_bounds(B, rank, world)
            D = lambda k: torch.from_numpy(batch[k][lo:hi].copy()).to(dev)
            outputs = {"heatmaps": D("heatmaps").requires_grad_(True), "offsets": D("offsets").requires_grad_(True),
                       "variances": D("variances").requires_grad_(True)}
            # on-the-fly target for even steps, target from HBM for odd ones
            target = None if step % 2 == 0 else D("target")
            weight = D("vis") if step % 2 == 0 else D("weight")
            out = loss_fn(outputs, target, weight, D("kps"), input_size=cfg.input_size)
            out["total_loss"].backward()
            results.append(({k: float(out[k]) for k in oc.LOSS_KEYS}, outputs["heatmaps"].grad.cpu().numpy(),
                            outputs["variances"].grad.cpu().numpy()))
        # the same steps again with both exchanges off the critical path: normalisers prefetched (here simply ahead of the
        # step, on a side stream), loss terms published and collected afterwards
        side = torch.cuda.Stream(device=dev)
        deferred = []
        for step in range(STEPS):
            batch = synth.make_batch(cfg, seed=50 + step, B=B)
            lo, hi = shard_bounds(B, rank, world)
            D = lambda k: torch.from_numpy(batch[k][lo:hi].copy()).to(dev)
            outputs = {"heatmaps": D("heatmaps").requires_grad_(True), "offsets": D("offsets").requires_grad_(True),
                       "variances": D("variances").requires_grad_(True)}
            target = None if step % 2 == 0 else D("target")
            weight = D("vis") if step % 2 == 0 else D("weight")
            kps = D("kps")
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                den = peer.denominators(loss_fn, weight, kps, (cfg.H, cfg.W), cfg.input_size, target_given=target is not None)
            torch.cuda.current_stream(dev).wait_stream(side)
            out = loss_fn(outputs, target, weight, kps, input_size=cfg.input_size, denominators=den, defer_losses=True)
            out["total_loss"].backward()
            glob = peer.collect_losses(dev)
            deferred.append((glob.cpu().numpy().copy(), outputs["heatmaps"].grad.cpu().numpy(), float(out["total_loss"])))
        torch.cuda.synchronize()
        q.put((rank, peer.timeouts(), results, deferred))
        dist.barrier()
        peer.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_exchange_through_peer_mailboxes():
    import infantposeestimation_gaussianbias_b200 as pkg
    pkg.load()
    from infantposeestimation_gaussianbias_b200 import FusionPoseLoss
    from infantposeestimation_gaussianbias_b200.sharded import shard_bounds
    world, B = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    # (a worker that dies puts nothing: fail fast instead of sitting out the queue's time-out)
    for _ in range(world):
        rank, timeouts, results, deferred = q.get(timeout=150)
        got[rank] = (timeouts, results, deferred)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    cfg = synth.CONFIGS["w32_256x192"]
    loss_fn = FusionPoseLoss(target_sigma=cfg.sigma)
    for step in range(STEPS):
        batch = synth.make_batch(cfg, seed=50 + step, B=B)
        D = lambda k: torch.from_numpy(batch[k]).cuda()
        outputs = {"heatmaps": D("heatmaps").requires_grad_(True), "offsets": D("offsets").requires_grad_(True),
                   "variances": D("variances").requires_grad_(True)}
        out = loss_fn(outputs, None if step % 2 == 0 else D("target"), D("vis") if step % 2 == 0 else D("weight"), D("kps"),
                      input_size=cfg.input_size)
        out["total_loss"].backward()
        gh, gv = outputs["heatmaps"].grad.cpu().numpy(), outputs["variances"].grad.cpu().numpy()
        for rank in range(world):
            timeouts, results, deferred = got[rank]
            assert timeouts == 0, f"rank {rank}: {timeouts} mailbox waits timed out"
            losses, rgh, rgv = results[step]
            # deferred mode: the collected losses are the global ones, the gradients the same bits, the value returned by the
            # step itself this rank's share only
            dl, dgh, share = deferred[step]
            np.testing.assert_allclose(dl, [float(out[k]) for k in oc.LOSS_KEYS], rtol=2e-6, atol=1e-9, err_msg=f"deferred step {step} rank {rank}")
            lo_, hi_ = shard_bounds(B, rank, world)
            assert np.array_equal(dgh, gh[lo_:hi_]), f"deferred step {step} rank {rank}: gradient differs"
            assert share < float(out["total_loss"]) * (1 + 1e-6)
            for k in oc.LOSS_KEYS:
                # the same per-tile numerators, added per rank first: last-bit differences only
                np.testing.assert_allclose(losses[k], float(out[k]), rtol=2e-6, atol=1e-9, err_msg=f"step {step} rank {rank} {k}")
            lo, hi = shard_bounds(B, rank, world)
            assert np.array_equal(rgh, gh[lo:hi]), f"step {step} rank {rank}: heatmap gradient differs from the global batch's"
            assert np.array_equal(rgv, gv[lo:hi])
    # both ranks hold the same bits (fixed rank order of the additions)
    for step in range(STEPS):
        assert got[0][1][step][0] == got[1][1][step][0]


def _late_peer_worker(rank, world, port, q):
    """rank 1 connects and then never calls: rank 0's kernels must give up after the time-out, poison the results and
    make the next call fail loudly (a late peer used to mean stale normalisers, silently)."""
    import time
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import infantposeestimation_gaussianbias_b200 as pkg
        pkg.load()
        from infantposeestimation_gaussianbias_b200 import _native as N
        from infantposeestimation_gaussianbias_b200.sharded import PeerExchange, ShardedFusionPoseLoss
        cfg = synth.CONFIGS["w32_256x192"]
        peer = PeerExchange(device=dev, timeout_s=1.0)
        if rank == 0:
            loss_fn = ShardedFusionPoseLoss(target_sigma=cfg.sigma, peer=peer)
            batch = synth.make_batch(cfg, seed=7, B=2)
            D = lambda k: torch.from_numpy(batch[k].copy()).to(dev)
            outputs = {"heatmaps": D("heatmaps"), "offsets": D("offsets"), "variances": D("variances")}
            t0 = time.time()
            out = loss_fn(outputs, None, D("vis"), D("kps"), input_size=cfg.input_size)
            total = float(out["total_loss"])                 # synchronises
            waited = time.time() - t0
            second = "no error"
            try:
                loss_fn(outputs, None, D("vis"), D("kps"), input_size=cfg.input_size)
            except N.GbcodecError as e:
                second = f"status {e.status}"
            q.put((total, waited, peer.timeouts(), second, N.lib().gbcodec_status_string(-7).decode()))
        dist.barrier()
        peer.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_a_late_peer_poisons_the_step_and_fails_the_next_call():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_late_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, waited, timeouts, second, name = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert np.isnan(total), "a step whose peers never answered must not return numbers"
    assert 0.9 <= waited < 60.0 and timeouts >= 1
    assert second == "status -7" and "peer" in name


def test_world_of_one_is_the_plain_step():
    """A peer context of a 1-rank job changes nothing."""
    import ctypes as C
    import infantposeestimation_gaussianbias_b200 as pkg
    pkg.load()
    from infantposeestimation_gaussianbias_b200 import _native as N, ops
    cfg = synth.CONFIGS["w32_256x192"]
    batch = synth.make_batch(cfg, seed=8, B=4)
    D = lambda k: torch.from_numpy(batch[k]).cuda()
    ctx, handle = C.c_void_p(), C.create_string_buffer(N.PEER_HANDLE_BYTES)
    N.check(N.lib().gbcodec_peer_create(0, 1, C.byref(ctx), handle), "peer_create")
    try:
        pairs = ops.pairs_flat(oc.COCO_SKELETON)
        alpha, fw = torch.tensor([0.5]).cuda(), torch.tensor([0.62]).cuda()
        args = (D("heatmaps"), D("offsets"), D("variances"), None, D("vis"), D("kps"), None, None, 192.0, 256.0,
                list(oc.DEFAULT_LAMBDAS), 2.0, 2.0, True, pairs, True, True, alpha, fw, 2, 3)
        a = ops.fusion_loss(*args)
        b = ops.fusion_loss(*args, int(ctx.value))
        for x, y in zip(a[:6], b[:6]):
            assert torch.equal(x, y)
        den = ops.loss_denominators(D("vis"), D("kps"), False, cfg.H, cfg.W, 192.0, 256.0, 2.0, pairs)
        assert torch.equal(b[6], den)
    finally:
        N.lib().gbcodec_peer_destroy(ctx)
