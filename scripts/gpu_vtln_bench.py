import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops
dev = torch.device("cuda", 0)
n, speakers, utts, T = 60, 109, 4, 1301
rows = speakers * utts * T
g = torch.Generator(device=dev).manual_seed(5)
x = torch.randn((rows, n), generator=g, device=dev)
alpha = (torch.rand(speakers, generator=g, device=dev) * 0.4 - 0.2).repeat_interleave(utts * T).contiguous()
gy = torch.randn((rows, n), generator=g, device=dev)
def timed(fn, steps=10, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
PEAK = 6452.8
for impl in ("cc", "tc"):
    ms = timed(lambda: ops.allpass_forward(x, alpha, n, impl=impl))
    print("fwd %s: %.4f ms  %.1f GB/s  %.3f of HBM peak" % (impl, ms, rows * (8 * n + 4) / ms / 1e6, rows * (8 * n + 4) / ms / 1e6 / PEAK))
if "bwd" in sys.argv:
    for impl in ("cc", "tc"):
        try:
            ms = timed(lambda: ops.allpass_backward(gy, x, alpha, n, impl=impl))
            print("bwd %s: %.4f ms  %.1f GB/s  %.3f of HBM peak" % (impl, ms, rows * (12 * n + 8) / ms / 1e6, rows * (12 * n + 8) / ms / 1e6 / PEAK))
        except TypeError as e:
            print("bwd", impl, "n/a", e)

if os.environ.get("B2W_LIB", "").find("vtfprof") >= 0:
    import ctypes
    from idiaptts_b200 import _lib
    bwd = os.environ.get("B2W_PROF_BWD") == "1"
    if bwd:
        ops.allpass_backward(gy, x, alpha, n, impl="tc")
        names = ["issuer: wait A_G", "issuer: wait matrix", "issuer: GEMM 1 issue", "issuer: wait A_X", "issuer: GEMM 2 issue", "-", "-", "-",
                 "group: wait gy rows", "group: wait A_G free", "group: convert G", "group: wait x rows + A_X free", "group: convert X",
                 "group: epilogue", "-", "-"]
    else:
        ops.allpass_forward(x, alpha, n, impl="tc")
        names = ["issuer: alpha runs", "issuer: wait matrix", "issuer: wait A", "issuer: MMA issue", "builder: build cycles (sum)", "builder: builds",
                 "builder: start of first build", "builder: fence + arrive (sum)", "group: wait raw tile", "group: read+convert+st",
                 "group: wait MMA", "group: wait out stage", "group: epilogue", "-", "-", "-"]
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    _lib.load().b2w_vtf_prof_read(ctypes.cast(buf, ctypes.c_void_p))
    for n_, v in zip(names, buf):
        print("  %-30s %10d cycles" % (n_, v))
