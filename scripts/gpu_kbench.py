"""Kernel micro-benchmark: times cheaptrick / mcep / d4c on a synthetic LJSpeech-shaped shard and prints checksums, so
kernel variants (B2W_LIB=variants/libb200world_<name>.so) can be compared for speed and for identical results.
usage: python scripts/gpu_kbench.py [--utts 2048] [--kernels cheaptrick,mcep,d4c] [--dump out.npz]"""
import argparse, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops, synthetic, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=1024)
ap.add_argument("--kernels", default="cheaptrick,mcep,d4c")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--dump", default=None)
a = ap.parse_args()
dev = torch.device("cuda", 0)
FS = 22050
waves, f0s = synthetic.make_corpus(a.utts, FS, seed=2, mean_dur=6.5, device=dev)
batch = ops.RaggedBatch.from_host([w.cpu().numpy() for w in waves], f0s, FS, device=dev)
F = batch.num_frames
chunk = min(F, 1 << 18)
status = ops.new_status(dev)
n_fft = ops.get_cheaptrick_fft_size(FS)
alpha = 0.455
ops.McepTables.get(59, alpha, n_fft, dev)
K = n_fft // 2 + 1
sp = torch.empty((chunk, (K + 7) // 8 * 8), dtype=torch.float32, device=dev)[:, :K]   # rows padded as in pipeline.WorldAnalyzer (aligned 16-byte loads)
mc = torch.empty((chunk, 60), dtype=torch.float32, device=dev)
print("lib %s, %d utts, %d frames, chunk %d" % (_lib.LIB_PATH, a.utts, F, chunk), flush=True)

def timeit(name, fn):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-11s %9.3f ms  %7.1f ns/frame" % (name, best, best * 1e6 / chunk), flush=True)
    return r

res = {}
ks = a.kernels.split(",")
ops.cheaptrick(batch, fft_size=n_fft, status=status, frame_lo=0, frame_hi=chunk, out=sp)
if "cheaptrick" in ks:
    timeit("cheaptrick", lambda: ops.cheaptrick(batch, fft_size=n_fft, status=status, frame_lo=0, frame_hi=chunk, out=sp))
    res["sp"] = sp[::97].cpu().numpy()
    print("  sp checksum %.12e" % sp.double().log().sum().item())
if "mcep" in ks:
    timeit("mcep", lambda: ops.mcep(sp, 59, alpha, is_power=True, out=mc.view(-1), out_stride=60, status=status))
    res["mc"] = mc[::97].cpu().numpy()
    print("  mc checksum %.12e" % mc.double().abs().sum().item())
if "d4c" in ks:
    coarse, voiced, _ = timeit("d4c", lambda: ops.d4c_coarse(batch, status=status, frame_lo=0, frame_hi=chunk))
    res["coarse"] = coarse.cpu().numpy(); res["voiced"] = voiced.cpu().numpy()
    v = voiced.bool()
    print("  d4c voiced %d, undecided %d, coarse checksum %.12e" % (int((voiced == 1).sum().item()), int((voiced == 2).sum().item()), coarse[v].double().sum().item()))
if "d4c_f64" in ks:
    c64, v64, _ = timeit("d4c_f64", lambda: ops.d4c_coarse(batch, status=status, frame_lo=0, frame_hi=chunk, precision="f64"))
    v = v64.bool()
    print("  d4c_f64 voiced %d, coarse checksum %.12e" % (int(v.sum().item()), c64[v].double().sum().item()))
    if "d4c" in ks:
        print("  fast vs f64: decisions differ on %d frames, max |coarse diff| %.3e dB" % (int((voiced != v64).sum().item()), (coarse[v] - c64[v]).abs().max().item()))
print("status", int(status.item()))
if a.dump:
    np.savez(a.dump, **res)

if os.environ.get("B2W_LIB", "").find("mcepprof") >= 0:
    import ctypes
    buf = (ctypes.c_longlong * 16)()
    _lib.load().b2w_mcep_prof_read(ctypes.cast(buf, ctypes.c_void_p))
    names = ["issuer: outside GEMM phase", "issuer: loop overhead", "issuer: wait stage", "issuer: wait D1 free", "issuer: GEMM1 issue",
             "issuer: wait A2", "issuer: GEMM2 issue", "issuer: wait GEMM2", "epi: loads (+ outside)", "epi: wait GEMM1",
             "epi: tmem ld + exp + split", "epi: wait A2 free", "epi: A2 stores + arrive", "-", "-", "-"]
    for n_, v in zip(names, buf):
        print("  %-28s %10d cycles" % (n_, v))
