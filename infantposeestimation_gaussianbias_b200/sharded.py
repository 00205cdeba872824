"""Batch-sharded six-term loss: one process per GPU, each rank holds B/R whole
images (so every limb pair stays on one rank).  The data path needs no
collective; the loss needs two tiny ones:

  1. before the kernel: all-reduce(sum) of [sum w, sum w_i*w_j] — the batch-global
     normalisers every term divides by (fusion_head.py:480,527,557,653,708,739);
  2. after it: all-reduce(sum) of the 7 local loss scalars (already divided by the
     global normalisers), for reporting.

With the global normalisers the local gradients ARE the global-batch gradients
restricted to the shard, so nothing else is exchanged.

Two transports for those 2 + 7 scalars:
  * NCCL all-reduces issued from here (default; also what the CPU/gloo tests exercise);
  * `PeerExchange`: the kernels that produce the scalars write them into every peer's mailbox over
    NVLink / NVSwitch peer memory and read the peers' values from their own (csrc/peer.cu,
    denoms_kernel / finalize_kernel in csrc/loss.cu) — no collective call on the step at all;
    torch.distributed is used once, to pass the 64-byte IPC handles around.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .fusion_head import LOSS_KEYS, FusionPoseLoss


def _default_denominators(loss: FusionPoseLoss, weight, gt_keypoints, target_given, H, W, input_size):
    from . import ops
    K = gt_keypoints.shape[1]
    sigma_enc = float(loss.encode_sigma if loss.encode_sigma is not None else loss.target_sigma)
    return ops.loss_denominators(weight.float(), gt_keypoints.float(), bool(target_given), H, W,
                                 float(input_size[0]), float(input_size[1]), sigma_enc,
                                 ops.pairs_flat(loss.pairs_for(K)))


class PeerExchange:
    """Mailboxes of a batch-sharded job (gbcodec_peer_* of include/gbcodec.h): one per rank, mapped into
    every other rank of `group` through CUDA IPC.  Needs one process per GPU on one NVLink / NVSwitch
    domain and an initialised process group (any backend) for the handle exchange."""

    def __init__(self, group=None, device: Optional[torch.device] = None, timeout_s: Optional[float] = None):
        """timeout_s: how long a kernel waits for a peer's scalars before it gives up (default 120 s: a rank that saves
        a checkpoint or runs a rank-local evaluation is late, not dead).  On expiry the step's normalisers / losses are
        NaN and the next sharded call on this rank raises GbcodecError (GBCODEC_ERR_PEER_TIMEOUT)."""
        import ctypes as C
        from . import _native as N
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > N.MAX_PEERS:
            raise RuntimeError(f"PeerExchange: at most {N.MAX_PEERS} ranks")
        self._lib = N.lib()
        self._ctx = C.c_void_p()
        handle = C.create_string_buffer(N.PEER_HANDLE_BYTES)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        error = None
        with torch.cuda.device(dev):
            try:
                N.check(self._lib.gbcodec_peer_create(self.rank, self.world, C.byref(self._ctx), handle), "peer_create")
            except Exception as e:           # every rank must still take part in the collectives below
                error = e
            gathered = [None] * self.world
            dist.all_gather_object(gathered, handle.raw if error is None else b"", group=group)
            if error is None and all(len(h) == N.PEER_HANDLE_BYTES for h in gathered):
                try:
                    N.check(self._lib.gbcodec_peer_connect(self._ctx, b"".join(gathered)), "peer_connect")
                except Exception as e:
                    error = e
            elif error is None:
                error = RuntimeError("PeerExchange: another rank could not create its mailbox")
            # one collective doubles as the barrier (nobody writes into a mailbox that is not mapped yet) and as the
            # vote: either every rank is connected or every rank raises
            ok = torch.tensor([0 if error is not None else 1], dtype=torch.int32, device=dev if dist.get_backend(group) == "nccl" else "cpu")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError(f"PeerExchange: peer-memory set-up failed on at least one rank"
                               f"{' (here: ' + str(error) + ')' if error is not None else ''}")
        if timeout_s is not None:
            self.set_timeout(timeout_s)

    def set_timeout(self, seconds: float) -> None:
        from . import _native as N
        N.check(self._lib.gbcodec_peer_set_timeout(self._ctx, float(seconds)), "peer_set_timeout")

    @property
    def address(self) -> int:
        return int(self._ctx.value or 0)

    def denominators(self, loss: FusionPoseLoss, target_weight: Tensor, gt_keypoints: Tensor, heatmap_hw, input_size,
                     target_given: bool = False, out: Optional[Tensor] = None) -> Tensor:
        """Global normaliser sums of a batch whose images are spread over the ranks (2 floats on the device), exchanged
        through the mailboxes on the CURRENT stream.  The sums depend on the visibility flags and the keypoints only, so
        a training loop calls this for batch t+1 on a side stream as soon as that batch is loaded, while step t runs,
        and hands the result to the loss as `denominators=`: the step then starts without waiting for anybody."""
        from . import ops
        H, W = heatmap_hw
        K = gt_keypoints.shape[1]
        sigma_enc = float(loss.encode_sigma if loss.encode_sigma is not None else loss.target_sigma)
        return ops.peer_denominators(target_weight.float(), gt_keypoints.float(), bool(target_given), int(H), int(W),
                                     float(input_size[0]), float(input_size[1]), sigma_enc, ops.pairs_flat(loss.pairs_for(K)),
                                     self.address, out)

    def collect_losses(self, device=None, steps_back: int = 0, out: Optional[Tensor] = None) -> Tensor:
        """The 7 global losses of the deferred step made `steps_back` (0..2) sharded calls ago, on the device."""
        from . import ops
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        return ops.peer_collect_losses(self.address, dev, steps_back, out)

    def timeouts(self) -> int:
        """Bounded spins that gave up so far on this rank (0 in a healthy job).  Synchronises the device."""
        import ctypes as C
        from . import _native as N
        n = C.c_int(0)
        N.check(self._lib.gbcodec_peer_status(self._ctx, C.byref(n)), "peer_status")
        return int(n.value)

    def close(self) -> None:
        if self._ctx:
            self._lib.gbcodec_peer_destroy(self._ctx)
            self._ctx = None


class ShardedFusionPoseLoss(FusionPoseLoss):
    """FusionPoseLoss for a rank that holds a shard of the global batch.

    `local_denominators` and `local_loss` exist so that the host-side logic can be
    exercised on CPU (gloo) with a stand-in for the CUDA ops; the defaults are the
    gbcodec ops."""

    def __init__(self, *args, process_group=None, peer: Optional[PeerExchange] = None,
                 local_denominators: Optional[Callable] = None, local_loss: Optional[Callable] = None, **kw):
        super().__init__(*args, **kw)
        self.process_group = process_group
        self.peer = peer
        self._local_denominators = local_denominators
        self._local_loss = local_loss

    def global_denominators(self, target_heatmaps, target_weight, gt_keypoints, heatmap_hw, input_size) -> Tensor:
        H, W = heatmap_hw
        given = target_heatmaps is not None and target_heatmaps.numel() > 0
        fn = self._local_denominators or (lambda *a: _default_denominators(self, *a))
        den = fn(target_weight, gt_keypoints, given, H, W, input_size).clone()
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(den, op=dist.ReduceOp.SUM, group=self.process_group)
        return den

    def forward(self, outputs: Dict[str, Tensor], target_heatmaps, target_weight, gt_keypoints,
                input_size: Tuple[int, int] = (192, 256), heatmap_size: Tuple[int, int] = (48, 64), **kw):
        if self.peer is not None:
            # both exchanges happen inside the kernels: the values returned are the global losses, the
            # gradients behind them this rank's share of the global-batch gradients
            return super().forward(outputs, target_heatmaps, target_weight, gt_keypoints, input_size, heatmap_size,
                                   peer=self.peer, **kw)
        H, W = outputs["heatmaps"].shape[-2:]
        den = self.global_denominators(target_heatmaps, target_weight, gt_keypoints, (H, W), input_size)
        if self._local_loss is not None:
            local = self._local_loss(outputs, target_heatmaps, target_weight, gt_keypoints, input_size, den)
        else:
            local = super().forward(outputs, target_heatmaps, target_weight, gt_keypoints, input_size, heatmap_size,
                                    denominators=den, **kw)
        stacked = torch.stack([local[k].detach() for k in LOSS_KEYS])
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(stacked, op=dist.ReduceOp.SUM, group=self.process_group)
        out = dict(local)
        for i, k in enumerate(LOSS_KEYS):
            # value = global loss; gradient = this rank's share of it
            out[k] = local[k] + (stacked[i] - local[k].detach())
        return out


def shard_bounds(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of B images over `world` ranks; the first B % world ranks get one more."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedCombinedLoss:
    """CombinedLoss (models/losses.py:205-290) for a rank that holds a shard of the global batch.  Its terms
    are plain means over the batch, so the only global quantity is the batch size: every rank runs the
    one-pass kernel with `norm_batch` = global batch (its gradients are then exactly the global-batch
    gradients restricted to the shard) and the five scalars are summed over the ranks for reporting.

    `local_loss(predictions, targets, norm_batch)` is the stand-in hook for the CPU (gloo) test."""

    def __init__(self, config, process_group=None, local_loss: Optional[Callable] = None):
        self.process_group = process_group
        self._local_loss = local_loss
        if local_loss is None:
            from .losses import CombinedLoss
            self.impl = CombinedLoss(config)

    def global_batch(self, local_batch: int, device) -> int:
        if not (dist.is_available() and dist.is_initialized()):
            return local_batch
        n = torch.tensor([local_batch], dtype=torch.int64, device=device)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=self.process_group)
        return int(n.item())

    def __call__(self, predictions: Dict[str, Tensor], targets: Dict[str, Tensor], global_batch: Optional[int] = None):
        some = next(iter(predictions.values()))
        if global_batch is None:
            global_batch = self.global_batch(some.shape[0], some.device)      # pass it in to avoid this host read
        if self._local_loss is not None:
            total, parts = self._local_loss(predictions, targets, global_batch)
        else:
            total, parts = self.impl(predictions, targets, norm_batch=global_batch)
        keys = sorted(parts)
        stacked = torch.stack([parts[k].detach() for k in keys])
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(stacked, op=dist.ReduceOp.SUM, group=self.process_group)
        out = {k: parts[k] + (stacked[i] - parts[k].detach()) for i, k in enumerate(keys)}    # value global, gradient local
        return out["total"], out
