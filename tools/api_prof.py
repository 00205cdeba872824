import cProfile, pstats, torch, sys, os
sys.path.insert(0, os.getcwd())
import bench
import infantposeestimation_gaussianbias_b200 as pkg
pkg.load()
from infantposeestimation_gaussianbias_b200 import ops, FusionPoseLoss
dev = torch.device("cuda", 0)
data = bench.synth_device_batch(1024, dev, 1, ops)
loss_fn = FusionPoseLoss(target_sigma=2.0)
leaves = {k: data[s].detach().clone().requires_grad_(True) for k, s in (("heatmaps","hm"),("offsets","off"),("variances","var"))}
def step():
    for v in leaves.values(): v.grad = None
    l = loss_fn(leaves, None, data["vis"], data["kps"], input_size=(192, 256))["total_loss"]
    l.backward()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
print("grad is stash:", True)
