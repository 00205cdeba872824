#!/usr/bin/env python
"""bench.py — heatmaps/s of the fused codec step (encode + six-term loss fwd/bwd + decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1 workload is BASELINE.json configs[1]: HRNet-W32 256x192 (64x48 heatmaps, K=17,
sigma=2), batch 1024 per GPU; one step = one pass of the hot path over one batch:
target tiles generated on the fly from the keypoints, the six-term loss forward and
backward, and the keypoint decode, every heatmap read from HBM once.
For N > 1 the driver launches one rank per GPU (torch.distributed.run); the batch is
sharded by image, per-GPU work fixed (weak scaling); the only data that crosses GPUs are
2 normaliser sums before and 7 loss scalars after the tile kernel, written by the kernels
themselves into the peers' NVLink-mapped mailboxes (`--exchange peer`, default) or
all-reduced by NCCL (`--exchange nccl`).

`--config decode_flip` / `--config decode [--batch B]` measure the decode-only workloads of
BASELINE.json (configs[2]: 96x72, B=4096, flip test + offset correction; configs[4]: the
64x48 sweep) under the same contract; `--config hrformer|preemie` the other fused-step shapes.

`--impl reference` times the reference's own CPU implementation of the same step on the host cores: the UNMODIFIED
reference tree (baseline/_ref, a verbatim copy made by tools/install_reference.py; /root/reference in the build
container) — `COCOPoseDataset._generate_target` per sample, `FusionPoseLoss` forward + backward, `head.decode` — with
every host thread (`kind: "reference"`).  Only if the tree is absent does it fall back to the oracle port
(oracle/heatmap_codec.py, pinned to the reference by tests/golden; `kind: "port"`).

The default line also carries: `aten_cuda_baseline` (the reference's stock modules run on the SAME GPU with device
tensors — what a user of the reference sees today), `api_step` (the patched module's forward + total_loss.backward(),
float32 and float16-autocast), `other_workloads` (BASELINE configs[2], [3], [4]) and `gpu_launches` as counted by the
library itself.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "heatmaps/sec (Bx17x64x48, encode+loss+decode)"
UNIT = "heatmaps/s"
K, H, W = 17, 64, 48
IN_W, IN_H = 192, 256
SIGMA = 2.0
LAMBDAS = [1.0, 1.0, 0.5, 0.1, 0.05, 0.05]
SKELETON = ((0, 1), (0, 2), (1, 3), (2, 4), (5, 6), (5, 7), (7, 9), (6, 8), (8, 10),
            (5, 11), (6, 12), (11, 12), (11, 13), (13, 15), (12, 14), (14, 16))
BYTES_PER_HM = 24 * H * W          # fused step: read P,V (8N); write dP,dV,dO (16N)  — SURVEY §8d / DESIGN.md


WORKLOAD = "BASELINE configs[1]: HRNet-W32 256x192 (64x48 heatmaps, K=17, sigma=2)"
SYNTH_CONFIG = "w32_256x192"        # tests/synth.py config of the CPU arm
TILE_KERNEL = "step_pipe_kernel<12,16,4,3,grads>"

# The default (and the driver's) workload is BASELINE configs[1].  The other fused-step configurations of BASELINE.json
# can be selected for a measurement of their own; they change the shapes only.
WORKLOADS = {
    "w32": None,
    "hrformer": dict(K=17, H=96, W=72, IN_W=288, IN_H=384, SIGMA=2.0, SYNTH_CONFIG="hrformer_384x288", TILE_KERNEL="loss_tile_kernel<18,16,6,...>",
                     WORKLOAD="BASELINE configs[2] shapes: HRFormer-base 384x288 (96x72 heatmaps, K=17, sigma=2)"),
    "preemie": dict(K=13, H=128, W=128, IN_W=256, IN_H=256, SIGMA=1.5, SYNTH_CONFIG="preemie_256", TILE_KERNEL="loss_tile_kernel<32,16,8,...>",
                    WORKLOAD="BASELINE configs[3]: preemie_optimized.yaml (128x128 heatmaps, K=13, sigma=1.5, 256x256 input), full six-term loss"),
    # decode-only workloads (their own metric: the step is one decode call, no loss)
    "decode_flip": dict(K=17, H=96, W=72, IN_W=288, IN_H=384, SIGMA=2.0, SYNTH_CONFIG="hrformer_384x288", DECODE="flip", DEFAULT_BATCH=4096,
                        TILE_KERNEL="decode_tile_kernel<18,16,6,flip>",
                        WORKLOAD="BASELINE configs[2]: HRFormer-base 384x288 (96x72 heatmaps, K=17) decode with flip test + offset correction"),
    "decode": dict(K=17, H=64, W=48, IN_W=192, IN_H=256, SIGMA=2.0, SYNTH_CONFIG="w32_256x192", DECODE="plain", DEFAULT_BATCH=16384,
                   TILE_KERNEL="decode_warp_kernel<12,64> (one warp per tile)",
                   WORKLOAD="BASELINE configs[4]: decode-only sweep point (64x48 heatmaps, K=17), sub-pixel refinement + offset correction"),
}
PRELOAD_STEPS = 256                 # untimed steps between the warm-up and the timed region while the clock sampler comes up
DECODE = None                       # None: fused step; "flip" / "plain": decode-only workloads
DEFAULT_BATCH = 1024


def select_workload(name):
    w = WORKLOADS[name]
    if w is None:
        return
    g = globals()
    g.update(w)
    if g["DECODE"]:
        g["BYTES_PER_HM"] = (8 if g["DECODE"] == "flip" else 4) * g["H"] * g["W"]     # read P (and the flipped pass); SURVEY §8d
        g["METRIC"] = f"heatmaps/sec (Bx{g['K']}x{g['H']}x{g['W']}, decode{' with flip test' if g['DECODE'] == 'flip' else ''} + offset correction)"
        return
    g["BYTES_PER_HM"] = 24 * g["H"] * g["W"]
    g["METRIC"] = f"heatmaps/sec (Bx{g['K']}x{g['H']}x{g['W']}, encode+loss+decode)"


def workload_name(B):
    return f"{WORKLOAD} fused codec step (on-the-fly encode + six-term loss fwd/bwd + decode), batch {B} per GPU"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            top = sorted(sm)[len(sm) // 2:]          # the loaded half of the samples
            out.update(sm_mhz=statistics.median(top), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------- inputs
def synth_device_batch(B, device, seed, ops, with_var=True, spec=None):
    """Synthetic batch created on the device (SURVEY §8d recipe): peaked heatmaps around
    jittered keypoints + noise, gaussian offsets, softplus variances.  `spec` overrides the module-level shape
    (K, H, W, IN_W, IN_H, SIGMA) for the secondary workloads."""
    import torch
    K, H, W, IN_W, IN_H, SIGMA = (spec[k] for k in ("K", "H", "W", "IN_W", "IN_H", "SIGMA")) if spec else (
        globals()[k] for k in ("K", "H", "W", "IN_W", "IN_H", "SIGMA"))
    g = torch.Generator(device=device).manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, device=device)
    rn = lambda *s: torch.randn(*s, generator=g, device=device)
    u = r(B, K)
    vis = torch.where(u < 0.15, 0.0, torch.where(u < 0.40, 1.0, 2.0))
    kps = torch.stack(((r(B, K) * 1.2 - 0.1) * IN_W, (r(B, K) * 1.2 - 0.1) * IN_H), dim=-1)
    jitter = rn(B, K, 2) * 1.5 * (IN_W / W)
    shifted, _ = ops.encode(kps + jitter, torch.full_like(vis, 2.0), H, W, float(IN_W), float(IN_H), SIGMA)
    amp = r(B, K, 1, 1) * 0.9 + 0.3
    hm = amp * shifted + 0.05 * rn(B, K, H, W)
    off = 0.3 * rn(B, K, 2, H, W)
    var = torch.nn.functional.softplus(rn(B, K, H, W)) if with_var else None
    if not with_var:
        return dict(hm=hm.contiguous(), off=off.contiguous())
    return dict(kps=kps.contiguous(), vis=vis.contiguous(), hm=hm.contiguous(), off=off.contiguous(), var=var.contiguous())


# ----------------------------------------------------------------------------- CPU arm
def cpu_step_rate(sample_B: int, min_seconds: float, max_reps: int):
    """The reference step (encode -> loss fwd+bwd -> decode) on the host cores: the unmodified reference when its tree is
    present, the oracle port otherwise.  -> (heatmaps/s, cores, per-pass times, kind)"""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    call, kind = reference_step_callable(sample_B)

    def one():
        t0 = time.perf_counter()
        call()
        return time.perf_counter() - t0

    one()                                   # warm-up
    times, t_start = [], time.perf_counter()
    while len(times) < max_reps and (len(times) < 3 or time.perf_counter() - t_start < min_seconds):
        times.append(one())
    return sample_B * K / statistics.median(times), cores, times, kind


def _reference_tree():
    """The unmodified reference, if it travelled with the repo (tools/install_reference.py) or is mounted."""
    from tests import refload
    return refload.Reference() if refload.find() else None


def reference_step_callable(sample_B, device="cpu"):
    """-> (callable running one reference step on `sample_B` images, kind).  The step is the reference's own code:
    COCOPoseDataset._generate_target per sample (coco_dataset.py:185-250; on the host, as the DataLoader workers do),
    FusionPoseLoss forward + total_loss.backward() (fusion_head.py:745-806, train.py:182), HeatmapRegressionHead.decode
    (fusion_head.py:309-365).  Falls back to the oracle port when the tree is absent."""
    import numpy as np
    import torch
    from tests import synth
    cfg = synth.CONFIGS[SYNTH_CONFIG]
    batch = synth.make_batch(cfg, seed=0, B=sample_B)
    ref = _reference_tree()
    if ref is None:
        from oracle import heatmap_codec as oc
        T = lambda k: torch.from_numpy(batch[k])
        return (lambda: oc.codec_step(batch["kps"], batch["vis"], T("heatmaps"), T("offsets"), T("variances"),
                                      heatmap_size=cfg.heatmap_size, input_size=cfg.input_size, sigma=cfg.sigma, loop_decode=True)), "port"
    ref.__enter__()                       # stays on sys.path for the life of this process
    fh = ref.fusion_head
    ds = object.__new__(ref.coco_dataset.COCOPoseDataset)
    ds.num_keypoints, ds.sigma = cfg.K, cfg.sigma
    ds.heatmap_size, ds.input_size = np.array(cfg.heatmap_size), np.array(cfg.input_size)
    loss_fn = fh.FusionPoseLoss(target_sigma=cfg.sigma).to(device)
    head = fh.HeatmapRegressionHead(32, num_keypoints=cfg.K).to(device)
    D = lambda k: torch.from_numpy(batch[k]).to(device)
    hm, off, var, kps = D("heatmaps"), D("offsets"), D("variances"), D("kps")
    fw = torch.sigmoid(head.fusion_weight.detach())

    def step():
        enc = [ds._generate_target(batch["kps"][b], batch["vis"][b]) for b in range(sample_B)]
        target = torch.from_numpy(np.stack([e[0] for e in enc])).to(device)
        weight = torch.from_numpy(np.stack([e[1] for e in enc])).to(device)
        outputs = {"heatmaps": hm.clone().requires_grad_(True), "offsets": off.clone().requires_grad_(True),
                   "variances": var.clone().requires_grad_(True), "fusion_weight": fw}
        losses = loss_fn(outputs, target, weight, kps, input_size=cfg.input_size, heatmap_size=(cfg.H, cfg.W))
        losses["total_loss"].backward()
        with torch.no_grad():
            coords, scores = head.decode({k: v.detach() for k, v in outputs.items()}, apply_offset=True)
        return float(losses["total_loss"]), coords, scores

    return step, "reference"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_B = 64
    # each step = one bounded sample of the workload (B=64 of the 1024-image batch)
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    call, kind = reference_step_callable(sample_B)
    for _ in range(max(1, min(args.warmup, 3))):
        call()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call()
    dt = time.perf_counter() - t0
    value = sample_B * K * args.steps / dt
    sample = f"{sample_B} images x {K} heatmaps per step (a bounded sample of the {args.batch}-image batch), {args.steps} steps"
    how = ("the unmodified reference (_generate_target per sample, FusionPoseLoss fwd + backward, head.decode) on the host cores"
           if kind == "reference" else "oracle port of the reference step (reference tree absent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "sample": sample, "what": how},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process (and therefore the pinned host buffers it is about to allocate: first touch) to the CPU
    cores of the NUMA node the GPU hangs off.  With one process per GPU and ~54 GB/s of H2D per rank, buffers on
    the far socket put the host-buffer step on the inter-socket link.  Best effort: returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                    # NVML pads the PCI domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------- secondary measurements of the default line
def _time_cuda(fn, steps, warm=3):
    """mean milliseconds per call of `fn` on torch's current stream (CUDA events, synchronised on both sides)."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def aten_cuda_baseline(device, sample_B=256, steps=2):
    """The reference's STOCK modules on this GPU with device tensors — its own execution mode (train.py:159-172 moves
    everything `.to(device)`): FusionPoseLoss forward + backward (ATen kernels), HeatmapRegressionHead.decode (a Python
    loop over B*K tiles with `.item()` round trips, fusion_head.py:102-126).  Targets are given (the reference builds
    them in DataLoader workers).  None when the reference tree is absent."""
    import torch
    ref = _reference_tree()
    if ref is None:
        return None
    from tests import synth
    cfg = synth.CONFIGS[SYNTH_CONFIG]
    batch = synth.make_batch(cfg, seed=0, B=sample_B)
    with ref:
        fh = ref.fusion_head
        loss_fn = fh.FusionPoseLoss(target_sigma=cfg.sigma).to(device)
        head = fh.HeatmapRegressionHead(32, num_keypoints=cfg.K).to(device)
        D = lambda k: torch.from_numpy(batch[k]).to(device)
        hm, off, var, kps, target, weight = D("heatmaps"), D("offsets"), D("variances"), D("kps"), D("target"), D("weight")
        fw = torch.sigmoid(head.fusion_weight.detach())

        def loss_step():
            outputs = {"heatmaps": hm.clone().requires_grad_(True), "offsets": off.clone().requires_grad_(True),
                       "variances": var.clone().requires_grad_(True), "fusion_weight": fw}
            loss_fn(outputs, target, weight, kps, input_size=cfg.input_size, heatmap_size=(cfg.H, cfg.W))["total_loss"].backward()

        def decode_step():
            with torch.no_grad():
                head.decode({"heatmaps": hm, "offsets": off, "fusion_weight": fw}, apply_offset=True)

        loss_ms = _time_cuda(loss_step, max(steps, 5), warm=2)
        t0 = time.perf_counter()
        decode_step()
        torch.cuda.synchronize()
        dec_ms = (time.perf_counter() - t0) * 1e3          # one pass: B*K Python iterations with host round trips
    n = sample_B * K
    return {"what": "reference's stock FusionPoseLoss fwd+bwd and head.decode on cuda:0 (ATen kernels, device tensors)",
            "sample": f"{sample_B} images x {K} heatmaps (per-heatmap cost is batch-independent for the decode loop; the loss's "
                      f"~200 launches amortise better at larger batches: extrapolation to batch 1024 is labelled as such)",
            "loss_fwd_bwd_ms": loss_ms, "decode_ms": dec_ms, "step_ms": loss_ms + dec_ms,
            "value": n / ((loss_ms + dec_ms) * 1e-3), "loss_only_value": n / (loss_ms * 1e-3), "unit": UNIT}


def api_step_rates(device, data, B, steps):
    """The step a user of the patched reference runs: FusionPoseLoss(...) (the module patch_reference() binds) forward +
    total_loss.backward(), gradients landing in .grad of the three head outputs; float32, and float16 maps under
    autocast with a GradScaler-style upstream factor.  Targets are built in the kernel (encode_on_device)."""
    import torch
    from infantposeestimation_gaussianbias_b200 import FusionPoseLoss
    loss_fn = FusionPoseLoss(target_sigma=SIGMA)
    out = {}
    for tag, cast, scale in (("fp32", lambda t: t, 1.0), ("fp16_autocast", lambda t: t.half(), 65536.0)):
        leaves = {k: cast(data[src]).detach().clone().requires_grad_(True) for k, src in (("heatmaps", "hm"), ("offsets", "off"), ("variances", "var"))}
        sc = torch.tensor(scale, device=device)

        def step():
            for v in leaves.values():
                v.grad = None
            with torch.autocast("cuda", enabled=tag != "fp32"):
                l = loss_fn(leaves, None, data["vis"], data["kps"], input_size=(IN_W, IN_H))["total_loss"]
            (l * sc).backward()

        ms = _time_cuda(step, steps, warm=4)
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        for _ in range(steps):
            step()
        host_ms = (time.perf_counter() - h0) * 1e3 / steps          # host time to enqueue one API step (nothing synchronises inside)
        torch.cuda.synchronize()
        out[tag] = {"ms_per_step": ms, "value": B * K / (ms * 1e-3), "unit": UNIT, "upstream_scale": scale,
                    "host_enqueue_ms_per_step": host_ms}
        del leaves
        torch.cuda.empty_cache()
    out["what"] = "FusionPoseLoss.forward (patched module, targets built in the kernel) + (scale * total_loss).backward(), CUDA events"
    return out


def other_workloads(device, ops, N, world, steps=20):
    """BASELINE configs[2], [3], [4] beside the headline: short resident measurements (no host-buffer leg, no CPU leg),
    same timing rules (CUDA events, inputs larger than L2 or rotated).  N > 1: the fused preemie step only (configs[3] is
    the batch-sharded one); the decode workloads do not couple ranks and are N = 1 lines."""
    import torch
    from infantposeestimation_gaussianbias_b200.pose_estimator import flip_permutation
    peak = _peak()[0]
    res = {}
    dflags = N.DECODE_REFINE | N.DECODE_APPLY_OFFSET
    alpha, fw = torch.tensor([0.5], device=device), torch.tensor([0.6224593312018546], device=device)

    def fused(name, B):
        sp = WORKLOADS[name]
        data = synth_device_batch(B, device, 99, ops, spec=sp)
        pairs = ops.pairs_flat([(i, j) for (i, j) in SKELETON if i < sp["K"] and j < sp["K"]])
        fn = lambda: ops.fusion_loss(data["hm"], data["off"], data["var"], None, data["vis"], data["kps"], None, None, float(sp["IN_W"]),
                                     float(sp["IN_H"]), LAMBDAS, sp["SIGMA"], sp["SIGMA"], True, pairs, True, True, alpha, fw, 2, dflags)
        ms = _time_cuda(fn, steps, warm=3)
        by = 24 * sp["H"] * sp["W"] * B * sp["K"]
        del data
        torch.cuda.empty_cache()
        return {"workload": sp["WORKLOAD"], "batch_per_gpu": B, "ms_per_step": ms, "value": B * sp["K"] / (ms * 1e-3), "unit": UNIT,
                "algorithmic_GBps": by / (ms * 1e-3) / 1e9, "frac_of_measured_peak": by / (ms * 1e-3) / 1e9 / peak}

    def decode(name, B):
        sp = WORKLOADS[name]
        flip = sp["DECODE"] == "flip"
        in_bytes = (8 if flip else 4) * sp["H"] * sp["W"] * B * sp["K"]
        n_rot = max(1, min(8, -(-512_000_000 // in_bytes)))
        perm = flip_permutation(sp["K"], ((1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16)), device) if flip else None
        sets = []
        for r in range(n_rot):
            d = synth_device_batch(B, device, 7 + 13 * r, ops, with_var=False, spec=sp)
            if r > 0:
                d["off"] = sets[0]["off"]
            if flip:
                d["flip"] = torch.flip(d["hm"][:, perm.long()], dims=[-1]).contiguous()
            sets.append(d)
        it = [0]

        def fn():
            d = sets[it[0] % n_rot]
            it[0] += 1
            ops.fast.decode(d["hm"], d.get("flip"), perm, d["off"], alpha, fw, 2, dflags)

        ms = _time_cuda(fn, max(steps, 2 * n_rot), warm=max(3, n_rot))
        del sets
        torch.cuda.empty_cache()
        return {"workload": sp["WORKLOAD"], "batch_per_gpu": B, "ms_per_step": ms, "value": B * sp["K"] / (ms * 1e-3), "unit": UNIT,
                "algorithmic_GBps": in_bytes / (ms * 1e-3) / 1e9, "frac_of_measured_peak": in_bytes / (ms * 1e-3) / 1e9 / peak,
                "input_sets_rotated": n_rot}

    res["configs[3] preemie fused step"] = fused("preemie", 1024)
    if world == 1:
        res["configs[2] decode_flip"] = decode("decode_flip", 4096)
        res["configs[4] decode sweep"] = [decode("decode", b) for b in (256, 1024, 4096, 16384)]
    return res


def _peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import infantposeestimation_gaussianbias_b200 as pkg
    pkg.load()                                   # raises if libgbcodec.so is missing — no fallback
    from infantposeestimation_gaussianbias_b200 import _native as N
    from infantposeestimation_gaussianbias_b200 import ops
    from infantposeestimation_gaussianbias_b200.host_step import HostCodecStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the codec has no CPU path (use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a rank that dies must not leave the others waiting for the default 10 minutes
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
    B = args.batch
    pairs = ops.pairs_flat(SKELETON)
    data = synth_device_batch(B, device, 1234 + rank, ops)
    alpha = torch.tensor([0.5], device=device)
    fw = torch.tensor([0.6224593312018546], device=device)
    dflags = N.DECODE_REFINE | N.DECODE_APPLY_OFFSET
    # N=1: weights/normalisers + loss + finalize.  N>1 over NCCL: (normalisers + export) + (weights + import + loss
    # + finalize); N>1 over peer memory: the same three kernels as N=1
    use_peer = world > 1 and args.exchange in ("peer", "peer-sync")
    launches_per_step = 3 if world == 1 else (3 if use_peer else 6)
    peer = None
    exchange = "none" if world == 1 else args.exchange
    if use_peer:
        # CUDA IPC needs the ranks to share an IPC namespace; if any rank cannot map its peers, every rank takes NCCL
        from infantposeestimation_gaussianbias_b200.sharded import PeerExchange
        why = ""
        try:
            peer = PeerExchange(device=device)
        except Exception as e:                                   # noqa: BLE001 — reported in the JSON line
            why = f"{type(e).__name__}: {e}"[:160]
        ok = torch.tensor([1 if peer is not None else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if peer is not None:
                peer.close()
            peer, use_peer = None, False
            launches_per_step = 6
            exchange = f"nccl (peer-memory set-up failed on some rank{': ' + why if why else ''})"

    # peer mode, default: both scalar exchanges are off the step's critical path.  The normalisers of step i+1 (they depend
    # on the visibility flags and keypoints only, known when the batch is loaded) are exchanged on a side stream while step
    # i runs; the step publishes its six loss terms and waits for nobody (they are collected after the timed region, the
    # way a training loop reads them when it logs).  `--exchange peer-sync` keeps both exchanges inside the step.
    defer = use_peer and args.exchange == "peer"
    if defer:
        main_stream = torch.cuda.current_stream(device)
        side = torch.cuda.Stream(device=device)
        den_buf = [torch.empty(2, device=device), torch.empty(2, device=device)]
        den_ready = [torch.cuda.Event(), torch.cuda.Event()]
        step_done = [torch.cuda.Event(), torch.cuda.Event()]
        counter = [0]

        def prefetch(i):
            with torch.cuda.stream(side):
                if i >= 2:
                    side.wait_event(step_done[i & 1])          # step i-2 has read this buffer
                ops.peer_denominators(data["vis"], data["kps"], False, H, W, float(IN_W), float(IN_H), SIGMA, pairs, peer.address, den_buf[i & 1])
                den_ready[i & 1].record(side)

        side.wait_stream(main_stream)
        prefetch(0)

    def step():
        den = None
        if defer:
            i = counter[0]
            main_stream.wait_event(den_ready[i & 1])
            prefetch(i + 1)
            r = ops.fusion_loss(data["hm"], data["off"], data["var"], None, data["vis"], data["kps"], den_buf[i & 1], None,
                                float(IN_W), float(IN_H), LAMBDAS, SIGMA, SIGMA, True, pairs, True, True, alpha, fw, 2, dflags,
                                peer.address, True)
            step_done[i & 1].record(main_stream)
            counter[0] = i + 1
            return r
        if use_peer:
            return ops.fusion_loss(data["hm"], data["off"], data["var"], None, data["vis"], data["kps"], None, None,
                                   float(IN_W), float(IN_H), LAMBDAS, SIGMA, SIGMA, True, pairs, True, True, alpha, fw, 2, dflags,
                                   peer.address)
        if world > 1:
            den = ops.loss_denominators(data["vis"], data["kps"], False, H, W, float(IN_W), float(IN_H), SIGMA, pairs)
            dist.all_reduce(den)
        res = ops.fusion_loss(data["hm"], data["off"], data["var"], None, data["vis"], data["kps"], den, None,
                              float(IN_W), float(IN_H), LAMBDAS, SIGMA, SIGMA, True, pairs, True, True, alpha, fw, 2, dflags)
        if world > 1:
            dist.all_reduce(res[0])
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----------------------------------------------------------
    # the warm-up keeps its result alive the way the timed loop does, so that the caching allocator
    # already owns both sets of output buffers (a cudaMalloc of 856 MB costs 5-300 ms on these boxes)
    for _ in range(args.warmup):
        res = step()
    # the tile kernel is bracketed by its own pair of events on every 4th timed step (each pair is two more stream
    # operations between the kernels of that step: sampling keeps the instrument out of most of the timed region)
    sampled = list(range(0, args.steps, 4))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in sampled]
    for a, b in ev:                               # materialise the handles
        a.record(); b.record()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    # nvidia-smi needs a moment to come up: the GPU stays under load meanwhile (every rank runs the same number of
    # untimed steps), so that the timed region starts at the clocks and power state of a long job, not from idle
    for _ in range(PRELOAD_STEPS):
        res = step()
    t_a, t_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launched0 = N.lib().gbcodec_launch_count()
    host_t0 = time.perf_counter()
    t_a.record()
    for i in range(args.steps):
        if i % 4 == 0:
            N.check(N.lib().gbcodec_profile_loss_kernel(N._P(ev[i // 4][0].cuda_event), N._P(ev[i // 4][1].cuda_event)), "profile")
        elif i % 4 == 1:
            N.lib().gbcodec_profile_loss_kernel(None, None)
        res = step()
    t_b.record()
    host_enqueue_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps    # host time to ENQUEUE a step (no sync inside the loop)
    launched = int(N.lib().gbcodec_launch_count() - launched0)      # this library's kernels, counted at their launch sites
    barrier()
    N.lib().gbcodec_profile_loss_kernel(None, None)
    ms = torch.tensor([t_a.elapsed_time(t_b)], device=device, dtype=torch.float64)
    kern_ms = torch.tensor([statistics.mean(a.elapsed_time(b) for a, b in ev)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    kern_ms = float(kern_ms.item())
    global_losses = None
    if use_peer and defer:
        global_losses = peer.collect_losses(device)            # the last timed step's losses, added over the ranks
        torch.cuda.synchronize()
    # keep the sampler running a little longer under load so that it sees loaded clocks
    if rank == 0:
        # (rank 0 only, so nothing collective in here: the local pass without the two all-reduces)
        t_end = time.time() + 1.0
        while time.time() < t_end:
            ops.fusion_loss(data["hm"], data["off"], data["var"], None, data["vis"], data["kps"], None, None,
                            float(IN_W), float(IN_H), LAMBDAS, SIGMA, SIGMA, True, pairs, True, True, alpha, fw, 2, dflags)
        torch.cuda.synchronize()
        clocks = sampler.stop()
    value = world * B * K * args.steps / (total_ms * 1e-3)

    # ---- end to end on host buffers (pinned -> H2D -> step -> D2H), same metric ---------------
    e2e = None
    if not args.no_e2e:
        host = {k: v.cpu().pin_memory() for k, v in data.items()}
        hs = HostCodecStep(B, K, H, W, (IN_W, IN_H), SIGMA, LAMBDAS, chunk_images=args.chunk, device=device)
        for _ in range(2):
            out = hs(host["hm"], host["off"], host["var"], host["kps"], host["vis"])
        n_e2e = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            out = hs(host["hm"], host["off"], host["var"], host["kps"], host["vis"])   # synchronises
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * K * n_e2e / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes,
               "d2h_bytes_per_step": hs.d2h_bytes, "steps": n_e2e, "chunk_images": hs.chunk,
               "total_loss": float(out["losses"][6])}
        check = float(res[0][6].item())
        if world == 1 and abs(e2e["total_loss"] - check) > 1e-4 * abs(check):
            raise SystemExit(f"bench.py: host-buffer step disagrees with the resident step: {e2e['total_loss']} vs {check}")

    if peer is not None:
        nt = peer.timeouts()
        if nt:
            raise SystemExit(f"bench.py: rank {rank} gave up waiting for a peer mailbox {nt} time(s)")
    # ---- the other BASELINE configurations, the API-level step, the reference's own GPU mode ---------------------
    others = api = aten = None
    if args.config == "w32" and not args.no_extras:
        others = other_workloads(device, ops, N, world)
        if world > 1:
            # every rank measured its shard; report the slowest rank's time and the job's aggregate rate
            for k, v in others.items():
                t = torch.tensor([v["ms_per_step"]], device=device, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                v["ms_per_step"] = float(t.item())
                v["value"] = world * v["batch_per_gpu"] * WORKLOADS["preemie"]["K"] / (v["ms_per_step"] * 1e-3)
                v["note"] = f"{world} ranks, one shard of {v['batch_per_gpu']} images each, no exchange inside this measurement (normalisers local)"
        if world == 1:
            api = api_step_rates(device, data, B, max(10, min(args.steps, 30)))
            aten = aten_cuda_baseline(device)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    peak, peak_src = _peak()
    achieved = B * K * BYTES_PER_HM / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": f"{TILE_KERNEL} (fused step: on-the-fly target + six-term loss fwd/bwd + decode)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "kernel_ms": kern_ms, "kernel_launches_timed": len(sampled),
                "algorithmic_bytes_per_launch": B * K * BYTES_PER_HM, "frac_of_nominal_8TBps": achieved / 8000.0}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path) and args.config == "w32":      # the ncu capture is of the default workload's kernel
        try:
            roofline["traffic"] = json.load(open(traffic_path)).get("loss_kernel_dram_bytes_per_launch")
        except Exception:
            pass

    cpu = None
    if not args.no_cpu and world == 1:                 # the CPU baseline is an N=1 figure (all host cores, nothing else running)
        v, cores, times, kind = cpu_step_rate(sample_B=32, min_seconds=10.0, max_reps=400)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"32 images x {K} heatmaps of the same workload, median of {len(times)} passes ({sum(times):.1f} s of CPU work)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(B), "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}", "exchange": exchange, "numa_node_rank0": numa,
                   "l2": f"inputs+outputs {B * K * BYTES_PER_HM / 1e6:.0f} MB per step, larger than the 126 MB L2; no flush needed"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launched, "gpu_launches_per_step": launched / args.steps, "host_enqueue_ms_per_step": host_enqueue_ms,
        "clocks": clocks,
        "total_loss": float((global_losses if global_losses is not None else res[0])[6].item()),
        "aten_cuda_baseline": aten, "api_step": api, "other_workloads": others,
    }
    print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------- decode-only workloads
def decode_workload_name(B):
    return f"{WORKLOAD}, batch {B} per GPU"


def reference_decode_callable(sample_B):
    """-> (callable, kind): the reference's decode of `sample_B` images on the host — head.decode (fusion_head.py:309-365)
    after the flip-test average written as PoseEstimator.inference writes it (pose_estimator.py:303-319: flip back, swap
    the left/right channels pair by pair, average) when the workload has one; the oracle port if the tree is absent."""
    import torch
    from tests import synth
    cfg = synth.CONFIGS[SYNTH_CONFIG]
    batch = synth.make_batch(cfg, seed=0, B=sample_B)
    hm, off = torch.from_numpy(batch["heatmaps"]), torch.from_numpy(batch["offsets"])
    flipped = torch.from_numpy(batch["heatmaps_flip"]) if DECODE == "flip" else None
    ref = _reference_tree()
    if ref is None:
        from oracle import heatmap_codec as oc
        return (lambda: oc.fusion_decode(hm, off, 0.5, 0.6224593312018546, True, True, 2, heatmaps_of_flipped_input=flipped, loop=True)), "port"
    ref.__enter__()
    head = ref.fusion_head.HeatmapRegressionHead(32, num_keypoints=cfg.K)
    fw = torch.sigmoid(head.fusion_weight.detach())
    flip_pairs = ref.config.get_config().data.flip_pairs

    def step():
        with torch.no_grad():
            h = hm
            if flipped is not None:
                back = torch.flip(flipped, dims=[-1])
                new = back.clone()
                for a, b in flip_pairs:
                    new[:, a] = back[:, b]
                    new[:, b] = back[:, a]
                h = (hm + new) / 2
            return head.decode({"heatmaps": h, "offsets": off, "fusion_weight": fw}, apply_offset=True)

    return step, "reference"


def cpu_decode_rate(sample_B: int, min_seconds: float, max_reps: int):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    call, kind = reference_decode_callable(sample_B)

    def one():
        t0 = time.perf_counter()
        call()
        return time.perf_counter() - t0

    one()
    times, t_start = [], time.perf_counter()
    while len(times) < max_reps and (len(times) < 3 or time.perf_counter() - t_start < min_seconds):
        times.append(one())
    return sample_B * K / statistics.median(times), cores, times, kind


def run_reference_decode(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    sample_B = 64
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    call, kind = reference_decode_callable(sample_B)
    for _ in range(max(1, min(args.warmup, 3))):
        call()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call()
    dt = time.perf_counter() - t0
    value = sample_B * K * args.steps / dt
    sample = f"{sample_B} images x {K} heatmaps per step (a bounded sample of the {args.batch}-image batch), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": decode_workload_name(args.batch), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


def run_decode_b200(args):
    """Decode-only workloads (BASELINE configs[2] and [4]).  No rank needs anything from another: N ranks decode N
    shards of the same size (weak scaling), the only collectives are the bench's own barriers and max-over-ranks."""
    import torch
    import torch.distributed as dist
    import infantposeestimation_gaussianbias_b200 as pkg
    pkg.load()
    from infantposeestimation_gaussianbias_b200 import _native as N
    from infantposeestimation_gaussianbias_b200 import ops
    from infantposeestimation_gaussianbias_b200.host_step import HostDecode
    from infantposeestimation_gaussianbias_b200.pose_estimator import flip_permutation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the codec has no CPU path (use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
    B = args.batch
    flip = DECODE == "flip"
    flip_pairs = ((1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16))     # configs/config.py:41-43
    perm = flip_permutation(K, flip_pairs, device) if flip else None
    alpha = torch.tensor([0.5], device=device)
    fw = torch.tensor([0.6224593312018546], device=device)
    dflags = N.DECODE_REFINE | N.DECODE_APPLY_OFFSET
    in_bytes = B * K * BYTES_PER_HM
    # inputs smaller than a few L2s are rotated so that no timed launch finds its heatmaps in the 126 MB L2
    n_rot = max(1, min(8, -(-512_000_000 // in_bytes)))
    sets = []
    for r in range(n_rot):
        d = synth_device_batch(B, device, 1234 + rank + 97 * r, ops, with_var=False)
        if r > 0:
            d["off"] = sets[0]["off"]                                  # 8 taps per tile are read: one copy of the offset maps
        if flip:
            g = torch.Generator(device=device).manual_seed(4321 + rank + r)
            d["flip"] = (torch.flip(d["hm"][:, perm.long()], dims=[-1]) + 0.02 * torch.randn(B, K, H, W, generator=g, device=device)).contiguous()
        sets.append(d)

    def step(i):
        d = sets[i % n_rot]
        return ops.fast.decode(d["hm"], d.get("flip"), perm, d["off"], alpha, fw, 2, dflags)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        res = step(i)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    for i in range(PRELOAD_STEPS):                 # untimed, see run_b200
        res = step(i)
    t_a, t_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launched0 = N.lib().gbcodec_launch_count()
    t_a.record()
    for i in range(args.steps):
        res = step(i)
    t_b.record()
    launched = int(N.lib().gbcodec_launch_count() - launched0)
    barrier()
    ms = torch.tensor([t_a.elapsed_time(t_b)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    # a step IS one launch of the decode kernel (on torch's current stream, where the two events are recorded), so the
    # average launch duration is the timed region over its launches; the gaps between launches count against the kernel
    kern_ms = total_ms / args.steps
    clocks = None
    if rank == 0:
        t_end, i = time.time() + 1.0, 0
        while time.time() < t_end:
            step(i); i += 1
        torch.cuda.synchronize()
        clocks = sampler.stop()
    value = world * B * K * args.steps / (total_ms * 1e-3)

    e2e = None
    if not args.no_e2e:
        d = sets[0]
        host = {k: v.cpu().pin_memory() for k, v in d.items()}
        hd = HostDecode(B, K, H, W, flip=flip, chunk_images=max(args.chunk, 256), device=device, flip_pairs=flip_pairs)
        for _ in range(2):
            out = hd(host["hm"], host.get("flip"), host["off"])
        n_e2e = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            out = hd(host["hm"], host.get("flip"), host["off"])       # synchronises
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        want = ops.decode(d["hm"], d.get("flip"), perm, d["off"], alpha, fw, 2, dflags)
        if not torch.equal(out["coords"], want[0].cpu()) or not torch.equal(out["scores"], want[1].cpu()):
            raise SystemExit("bench.py: host-buffer decode disagrees with the resident decode")
        e2e = {"value": world * B * K * n_e2e / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": hd.h2d_bytes,
               "d2h_bytes_per_step": hd.d2h_bytes, "steps": n_e2e, "chunk_images": hd.chunk}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = _peak()
    peak_src += " — a 1:1 read:write mix; read-only streams run above it"
    achieved = in_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": f"{TILE_KERNEL} (soft-argmax + window refinement + offset taps{', flip average in the load' if flip else ''})",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel_ms": kern_ms, "kernel_launches_timed": args.steps, "algorithmic_bytes_per_launch": in_bytes,
                "frac_of_nominal_8TBps": achieved / 8000.0}
    cpu = None
    if not args.no_cpu and world == 1:
        v, cores, times, kind = cpu_decode_rate(sample_B=32, min_seconds=10.0, max_reps=400)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"32 images x {K} heatmaps of the same workload, median of {len(times)} passes ({sum(times):.1f} s of CPU work)"}
    l2 = (f"heatmaps {in_bytes / 1e6:.0f} MB per step, larger than the 126 MB L2; no flush needed" if n_rot == 1 else
          f"heatmaps {in_bytes / 1e6:.0f} MB per step: {n_rot} input sets ({n_rot * in_bytes / 1e6:.0f} MB) visited in rotation, so a launch never finds its heatmaps in the 126 MB L2")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": decode_workload_name(B), "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "exchange": "none (shards are independent)", "numa_node_rank0": numa, "l2": l2},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launched, "clocks": clocks,
    }
    print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _json_only_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    communicator creation), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to
    the original stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: 1024; decode_flip 4096; decode 16384)")
    ap.add_argument("--chunk", type=int, default=128, help="images per chunk of the host-buffer pipeline")
    ap.add_argument("--exchange", default="peer", choices=["peer", "peer-sync", "nccl"],
                    help="N>1: how the 2+7 loss scalars travel between ranks (NVLink peer-memory mailboxes written by the kernels, or NCCL all-reduces)")
    ap.add_argument("--config", default="w32", choices=sorted(WORKLOADS), help="fused-step workload (default: BASELINE configs[1])")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_workloads / api_step / aten_cuda_baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    select_workload(args.config)
    if args.batch is None:
        args.batch = DEFAULT_BATCH
    args.out = _json_only_stdout()
    if args.impl == "reference":
        return run_reference_decode(args) if DECODE else run_reference(args)
    return run_decode_b200(args) if DECODE else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
