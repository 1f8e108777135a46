import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes
from build_umma_test import build_umma_test
lib = ctypes.CDLL(build_umma_test())  # development library, NOT part of libb200world.so
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for (n, k) in ((16, 8), (32, 8), (16, 16), (32, 64), (128, 32), (128, 64), (64, 16), (256, 16)):
    a = torch.randn(128, k, device=dev)
    bt = torch.randn(n, k, device=dev) * torch.logspace(-3, 0, k, device=dev)[None, :]
    ws = torch.zeros(2 * n * k, device=dev)
    d = torch.full((128, n), float("nan"), device=dev)
    rc = lib.b2w_test_umma_gemm(a.data_ptr(), bt.data_ptr(), n, k, ws.data_ptr(), d.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = a.double() @ bt.double().T
    err = (d.double() - ref).abs().max().item()
    tf32 = (a.double() @ bt.double().T - (a @ bt.T).double()).abs().max().item()
    print("N=%3d K=%2d rc=%d max abs err %.3e (ref max %.2f, torch fp32 matmul err %.1e) nan=%d" % (n, k, rc, err, ref.abs().max().item(), tf32, int(torch.isnan(d).sum())))
    if err > 1e-3:
        # diagnostics: which rows / cols are right
        dd = (d.double() - ref).abs()
        print("   rows ok:", (dd.max(1).values < 1e-4).sum().item(), "/128  cols ok:", (dd.max(0).values < 1e-4).sum().item(), "/", n)
        print("   d[0,:8]  ", d[0, :8].tolist()); print("   ref[0,:8]", ref[0, :8].tolist())
