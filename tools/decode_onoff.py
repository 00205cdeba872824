"""Is the scalar warp the step kernel's critical path?  Time the fused step with and without the keypoint decode
(the decode lives in the scalar warp only: three dependent global round trips per tile)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import infantposeestimation_gaussianbias_b200 as pkg
pkg.load()
from infantposeestimation_gaussianbias_b200 import ops, _native as N
dev = torch.device("cuda", 0)
d = bench.synth_device_batch(1024, dev, 1, ops)
pairs = ops.pairs_flat(bench.SKELETON)
a, f = torch.tensor([0.5], device=dev), torch.tensor([0.62], device=dev)
for dec in (True, False, True, False):
    fn = lambda: ops.fusion_loss(d["hm"], d["off"], d["var"], None, d["vis"], d["kps"], None, None, 192.0, 256.0, bench.LAMBDAS, 2.0, 2.0, True, pairs,
                                 True, dec, a, f, 2, 3)
    evs = []
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ks = []
    for i in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e1.record()          # materialise the handles
        N.check(N.lib().gbcodec_profile_loss_kernel(N._P(e0.cuda_event), N._P(e1.cuda_event)), "p")
        fn(); evs.append((e0, e1))
    N.lib().gbcodec_profile_loss_kernel(None, None)
    torch.cuda.synchronize()
    print("decode", dec, "kernel ms", sum(x.elapsed_time(y) for x, y in evs) / len(evs))
