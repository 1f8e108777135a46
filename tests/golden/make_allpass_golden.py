"""Builds tests/golden/allpasswarp_golden.npz by running the reference's own AllPassWarp layer (CPU torch) here.

The reference layer (idiaptts/src/neural_networks/pytorch/layers/AllPassWarp.py) is importable in this container; its
float32 polynomial tensor is finite and accurate only for small warp_matrix_size, so the
vectors use n = 8 and n = 20 with |alpha| <= 0.15.  Recorded: inputs, alphas (two stacked warping layers, which the
layer combines), outputs, and autograd gradients w.r.t. input and alphas for a fixed upstream gradient."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from idiaptts.src.neural_networks.pytorch.layers.AllPassWarp import AllPassWarp  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "allpasswarp_golden.npz")


def main():
    out = {}
    rng = np.random.default_rng(5)
    for n in (8, 20):
        T, B, blocks = 7, 3, 3
        layer = AllPassWarp(n).double()
        layer.w_matrix_3d = layer.w_matrix_3d.double()
        x = torch.tensor(rng.standard_normal((T, B, blocks * n)), dtype=torch.float64, requires_grad=True)
        a1 = torch.tensor(rng.uniform(-0.1, 0.1, (T, B, 1)), dtype=torch.float64, requires_grad=True)
        a2 = torch.tensor(rng.uniform(-0.05, 0.05, (T, B, 1)), dtype=torch.float64, requires_grad=True)
        gy = torch.tensor(rng.standard_normal((T, B, blocks * n)), dtype=torch.float64)
        y, combined = layer(x.clone(), [a1, a2])
        (y * gy).sum().backward()
        p = "n%d/" % n
        out[p + "x"] = x.detach().numpy()
        out[p + "a1"] = a1.detach().numpy()
        out[p + "a2"] = a2.detach().numpy()
        out[p + "gy"] = gy.numpy()
        out[p + "y"] = y.detach().numpy()
        out[p + "combined"] = combined.detach().numpy()
        out[p + "gx"] = x.grad.numpy()
        out[p + "ga1"] = a1.grad.numpy()
        out[p + "ga2"] = a2.grad.numpy()
        print(n, float(y.abs().max()), float(x.grad.abs().max()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
