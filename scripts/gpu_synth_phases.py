"""Where does a batched synthesis call spend its time?  (scripts/, diagnostics)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops, pipeline, synthetic
dev = torch.device("cuda", 0)
fs, U = 22050, 256
waves, f0s = synthetic.make_corpus(U, fs, seed=4, mean_dur=6.5, device=dev)
batch = ops.RaggedBatch.from_host([w.cpu().numpy() for w in waves], f0s, fs, device=dev)
an = pipeline.WorldAnalyzer(fs, 60, device=dev)
feats, _, _ = an.extract(batch)
syn = pipeline.WorldSynthesizer(fs, 60, device=dev)
D = 60
def sync():
    torch.cuda.synchronize()
    return time.perf_counter()
for it in range(3):
    y, out_off, st = syn.synthesize(feats, batch.frame_off)
sync()
for it in range(2):
    t0 = sync()
    y, out_off, st = syn.synthesize(feats, batch.frame_off)
    t1 = sync()
    print("overlapped total %.2f ms" % ((t1 - t0) * 1e3))
# sequential, phase by phase (synchronised between phases: host + device time of each)
for it in range(2):
    t0 = sync()
    lf0 = feats[:, D].double(); vuv = (feats[:, D + 1] >= 0.5); f0 = torch.exp(lf0)
    vuv = vuv & ~(f0 < 30); f0 = torch.where(vuv, f0, torch.zeros_like(f0)).contiguous()
    bap = feats[:, D + 2:].double().contiguous()
    t1 = sync()
    pow_sp = ops.mc2sp(feats, syn.alpha, syn.n_fft, scale=1.0, do_exp=True, out_dtype=torch.float64, order=D - 1, mc_stride=feats.shape[1], square=True)
    ap = ops.decode_aperiodicity(bap, fs, syn.n_fft)
    t2 = sync()
    h0 = time.perf_counter()
    plan = ops.synth_timebase(f0, batch.frame_off, fs, syn.n_fft, 5.0)
    h1 = time.perf_counter()
    t3 = sync()
    y, out_off, st = ops.synth_render(plan, pow_sp, ap, out_dtype=torch.float32)
    h2 = time.perf_counter()
    t4 = sync()
    print("sequential: prep %.2f planes %.2f timebase %.2f (host part %.2f) render+ola %.2f (host part %.2f) total %.2f ms" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (h1 - h0) * 1e3, (t4 - t3) * 1e3, (h2 - t3) * 1e3, (t4 - t0) * 1e3))
# sequential without intermediate syncs
for it in range(2):
    t0 = sync()
    lf0 = feats[:, D].double(); vuv = (feats[:, D + 1] >= 0.5); f0 = torch.exp(lf0)
    vuv = vuv & ~(f0 < 30); f0 = torch.where(vuv, f0, torch.zeros_like(f0)).contiguous()
    bap = feats[:, D + 2:].double().contiguous()
    pow_sp = ops.mc2sp(feats, syn.alpha, syn.n_fft, scale=1.0, do_exp=True, out_dtype=torch.float64, order=D - 1, mc_stride=feats.shape[1], square=True)
    ap = ops.decode_aperiodicity(bap, fs, syn.n_fft)
    y, out_off, st = ops.synthesize(f0, pow_sp, ap, batch.frame_off, fs, 5.0, out_dtype=torch.float32)
    t4 = sync()
    print("sequential, no intermediate syncs: total %.2f ms" % ((t4 - t0) * 1e3))
