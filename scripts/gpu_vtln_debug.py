import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops
from oracle import glue_np
dev = torch.device("cuda", 0)
n, blocks = 60, 1
rng = np.random.default_rng(261)
runs = [300, 128, 5, 511, 77, 640]
rows = sum(runs)
x = rng.standard_normal((rows, n)).astype(np.float32)
gy = rng.standard_normal((rows, n)).astype(np.float32)
al = np.concatenate([np.full(r, rng.uniform(-0.2, 0.2), np.float32) for r in runs])
xd, ad, gd = torch.from_numpy(x).to(dev), torch.from_numpy(al).to(dev), torch.from_numpy(gy).to(dev)
mean = rng.standard_normal(n).astype(np.float32); std = rng.uniform(0.5, 2.0, n).astype(np.float32)
md, sd = torch.from_numpy(mean).to(dev), torch.from_numpy(std).to(dev)
gx_ref, ga_ref = glue_np.allpass_warp_backward(gy / std, x.astype(np.float64) * std + mean, al, n)
gx_ref = gx_ref * std
for impl in ("cc", "tc"):
    gx, ga = ops.allpass_backward(gd, xd, ad, n, md, sd, impl=impl)
    gx, ga = gx.cpu().numpy(), ga.cpu().numpy()
    ex = np.abs(gx - gx_ref).max(1)
    ea = np.abs(ga - ga_ref) / (np.abs(ga_ref) + 1.0)
    print(impl, "gx max err %.3e at row %d; ga max rel err %.3e at row %d" % (ex.max(), ex.argmax(), ea.max(), ea.argmax()))
    bad = np.where(ea > 1e-4)[0]
    print("   bad ga rows:", bad[:10], len(bad), "ga", ga[bad[:4]], "ref", ga_ref[bad[:4]])
    badx = np.where(ex > 5e-5)[0]
    print("   bad gx rows:", badx[:10], len(badx))
