import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes
from build_umma_test import build_umma_test
lib = ctypes.CDLL(build_umma_test())  # development library, NOT part of libb200world.so
dev = torch.device("cuda", 0)
out = torch.zeros(4, dtype=torch.int64, device=dev)
for m, f16 in ((128, 0), (64, 0), (128, 1), (64, 1)):
    for n in (32, 64, 128):
        for count, nacc in ((1, 1), (24, 1), (96, 1)):
            rc = lib.b2w_probe_umma(n, count, nacc, m, f16, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            o = out.cpu().tolist()
            print("M=%3d %s N=%3d count=%3d nacc=%d: cycles %s  -> %.1f cycles/MMA (last rep)" % (m, "f16 " if f16 else "tf32", n, count, nacc, o, o[-1] / count), flush=True)
