import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops, synthetic
dev = torch.device("cuda", 0)
FS = 22050
waves, f0s = synthetic.make_corpus(128, FS, seed=2, mean_dur=6.5, device=dev)
batch = ops.RaggedBatch.from_host([w.cpu().numpy() for w in waves], f0s, FS, device=dev)
F = batch.num_frames
n_fft = ops.get_cheaptrick_fft_size(FS)
sp, st = ops.cheaptrick(batch, fft_size=n_fft, out_dtype=torch.float32)
it = torch.zeros(F, dtype=torch.int32, device=dev)
mc, st = ops.mcep(sp, 59, 0.455, is_power=True, iters=it)
it = it.cpu().numpy()
print("frames", F, "mean passes", it.mean(), "hist", np.bincount(it)[:32].tolist())
T = F // 128
tm = it[:T * 128].reshape(T, 128).max(1)
print("per-tile max: mean %.2f  hist %s" % (tm.mean(), np.bincount(tm)[:32].tolist()))
act = np.array([(it[:T * 128].reshape(T, 128) > p).sum(1).mean() for p in range(1, 16)])
print("mean active frames per tile entering solve of pass p=1..15:", np.round(act, 1).tolist())
f0 = np.concatenate(f0s)
print("mean passes voiced %.2f unvoiced %.2f" % (it[f0 > 0].mean(), it[f0 == 0].mean()))
