import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops
from oracle import world_np
dev = torch.device("cuda", 0)
fs, T = 16000, 400
f0 = np.zeros(T); f0[150:300] = np.linspace(110.0, 180.0, 150)
rng = np.random.default_rng(0)
sp = np.abs(rng.standard_normal((T, 513))) * 1e-3 + 1e-4
ap = np.clip(rng.uniform(0.05, 0.9, (T, 513)), 0.001, 0.999)
ylen = int(T * 5.0 * fs / 1000)
idx, shift, ivuv = world_np.synthesis_time_base(f0, fs, 0.005, ylen, 1024)
dbg = {}
foff = torch.tensor([0, T], dtype=torch.int64, device=dev)
y, _, st = ops.synthesize(torch.from_numpy(f0).to(dev), torch.from_numpy(sp).to(dev), torch.from_numpy(ap).to(dev), foff, fs, debug=dbg)
P = int(dbg["num_pulses"][0]); gi = dbg["pulse_index"][:P].cpu().numpy(); gs = dbg["pulse_shift"][:P].cpu().numpy(); gv = dbg["pulse_vuv"][:P].cpu().numpy()
print("pulses oracle/gpu", len(idx), P)
n = min(len(idx), P)
bad = np.nonzero(gi[:n] != idx[:n])[0]
print("index mismatches", len(bad), bad[:10], gi[bad[:10]], idx[bad[:10]])
print("shift max abs diff (matching)", np.abs(gs[:n] - shift[:n])[gi[:n] == idx[:n]].max())
print("vuv mismatch", (gv[:n] != (ivuv[idx[:n]] > 0.5)).sum())
y_ref = world_np.synthesize(f0, sp, ap, fs)
yy = y.cpu().numpy()
err = np.abs(yy - y_ref)
print("max err", err.max(), "at", err.argmax(), "first err>1e-9 at", np.nonzero(err > 1e-9)[0][:5], "count", (err > 1e-9).sum())
# per-pulse response check against oracle responses: recompute the oracle response of the worst region's pulses
w = err.argmax()
near = np.nonzero(np.abs(idx - w) < 600)[0]
print("pulses near worst", near[:5], idx[near[:5]], "vuv", ivuv[idx[near[:5]]], "noise sizes", np.diff(idx)[near[:5]])
resp = []
world_np.synthesize(f0, sp, ap, fs, responses=resp)
R = dbg["response"][:P].cpu().numpy()
errs = np.array([np.abs(R[p] - resp[p][2]).max() for p in range(P)])
print("per-pulse max err: unvoiced", errs[gv == 0].max(), "voiced", errs[gv == 1].max(), "n voiced", (gv == 1).sum())
p = int(np.argmax(errs))
d = R[p] - resp[p][2]
print("worst pulse", p, "idx", idx[p], "err by quarter", [np.abs(d[i * 256:(i + 1) * 256]).max() for i in range(4)], "resp max", np.abs(resp[p][2]).max())
per, aper, _ = resp[p]
ns = idx[min(P - 1, p + 1)] - idx[p]
print("periodic part max", np.abs(per).max() * np.sqrt(ns) / 1024, "aperiodic part max", np.abs(aper).max() / 1024)
# is the error proportional to the periodic or aperiodic oracle part?
print("corr with periodic", np.corrcoef(d, per)[0, 1], "corr with aperiodic", np.corrcoef(d, aper)[0, 1])
print("err energy / periodic energy", (d ** 2).sum() / ((per * np.sqrt(ns) / 1024) ** 2).sum(), " / aperiodic", (d ** 2).sum() / ((aper / 1024) ** 2).sum())
print("shift", shift[p], "coef*512", 2 * np.pi * shift[p] * fs / 1024 * 512)
