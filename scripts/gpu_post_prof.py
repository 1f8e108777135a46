"""Driver for the ncu capture of the post-network operators (deltas, MLPG, metrics) at the postprocess256 workload's size."""
import numpy as np
import torch

from idiaptts_b200 import ops

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
nu, D, W = 256, 60, 64
lengths = rng.integers(1100, 1500, nu)
fo_np = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
F = int(fo_np[-1])
fo = torch.from_numpy(fo_np).to(dev)
fu = torch.from_numpy(np.repeat(np.arange(nu, dtype=np.int32), lengths)).to(dev)
sub = torch.randn((F, W), device=dev)
for _ in range(3):
    d, dd = ops.deltas(sub, fo)
    tri = torch.cat((sub[:, :D], d[:, :D], dd[:, :D]), dim=1).contiguous()
    x = ops.mlpg(tri, torch.ones(3 * D, dtype=torch.float64, device=dev), fo, D)
    acc = ops.world_metrics(sub, sub + 0.1, fu, nu, D, 2)
torch.cuda.synchronize()
print("ok", F, float(x.abs().mean()), float(acc.sum()))
