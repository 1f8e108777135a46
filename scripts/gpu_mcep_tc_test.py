import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops
from oracle import glue_np, sptk_np, world_np
g = np.load(os.path.join(ROOT, "tests", "golden", "ljspeech_world_golden.npz"))
dev = torch.device("cuda", 0)
ID = "LJ001-0008"
x = g[ID + "/wav"].astype(np.float64) / 32768.0
x = np.append(x[0], x[1:] - 0.97 * x[:-1])
c = g[ID + "/cmp"]; T = c.shape[0]
f0 = np.where(c[:, 63] > 0, np.exp(c[:, 60].astype(np.float64)), 0.0)
sp = world_np.cheaptrick(x, f0, world_np.temporal_positions(T), 16000)
for order, alpha in ((19, 0.58), (59, 0.58), (59, 0.41)):
    ref = [sptk_np.mcep_frame(np.sqrt(a), order, alpha, eps=1e-8) for a in sp]
    rm = np.stack([r[0] for r in ref]); rit = np.array([r[1] for r in ref])
    plane = torch.from_numpy(sp.astype(np.float32)).to(dev)
    for impl in ("cc", "tc"):
        it = torch.zeros(T, dtype=torch.int32, device=dev)
        mc, st = ops.mcep(plane, order, alpha, is_power=True, out_dtype=torch.float64, iters=it, impl=impl)
        torch.cuda.synchronize()
        mc = mc.cpu().numpy()
        print("order %d alpha %.2f %s: max|d| %.3e MCD %.3e iters equal %.3f status %d nan %d" % (
            order, alpha, impl, np.nanmax(np.abs(mc - rm)), glue_np.mcd_db(rm, mc), np.mean(it.cpu().numpy() == rit), int(st.item()), int(np.isnan(mc).sum())), flush=True)
# timing on a big random-ish batch: tile the utterance
big = torch.from_numpy(np.tile(sp.astype(np.float32), (700, 1))).to(dev)
for impl in ("cc", "tc"):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        mc, st = ops.mcep(big, 59, 0.455, is_power=True, impl=impl)
        torch.cuda.synchronize(); dt = time.time() - t0
    print("%s: %d frames in %.1f ms -> %.1f ns/frame" % (impl, big.shape[0], dt * 1e3, dt / big.shape[0] * 1e9), flush=True)
