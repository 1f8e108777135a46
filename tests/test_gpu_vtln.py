"""Neural-VTLN all-pass warp kernels: against the reference layer's own outputs and autograd gradients (golden vectors made
by tests/golden/make_allpass_golden.py from the reference's AllPassWarp on CPU) and against the fp64 oracle at n = 60."""
import os

import numpy as np
import pytest
import torch

from oracle import glue_np

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [8, 20])
def test_layer_matches_reference_goldens(n):
    from idiaptts_b200.AllPassWarp import AllPassWarp
    g = np.load(os.path.join(ROOT, "tests", "golden", "allpasswarp_golden.npz"))
    p = "n%d/" % n
    dev = torch.device("cuda", 0)
    x = torch.tensor(g[p + "x"], dtype=torch.float32, device=dev, requires_grad=True)
    a1 = torch.tensor(g[p + "a1"], dtype=torch.float32, device=dev, requires_grad=True)
    a2 = torch.tensor(g[p + "a2"], dtype=torch.float32, device=dev, requires_grad=True)
    gy = torch.tensor(g[p + "gy"], dtype=torch.float32, device=dev)
    layer = AllPassWarp(n).to(dev)
    x_before = x.detach().clone()
    y, combined = layer(x, [a1, a2])
    assert torch.equal(x.detach(), x_before)  # input not modified in place (reference test_AllPassLayer.py)
    (y * gy).sum().backward()
    np.testing.assert_allclose(combined.detach().cpu().numpy(), g[p + "combined"], atol=1e-7)
    np.testing.assert_allclose(y.detach().cpu().numpy(), g[p + "y"], atol=5e-6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[p + "gx"], atol=5e-6)
    np.testing.assert_allclose(a1.grad.cpu().numpy(), g[p + "ga1"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(a2.grad.cpu().numpy(), g[p + "ga2"], atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("n,blocks", [(60, 1), (60, 3), (30, 3), (5, 2)])
def test_kernels_vs_fp64_oracle(n, blocks):
    """n = 60 is where the reference's float32 polynomial tensor overflows to inf/NaN; the fp64 freqt recursion is the oracle."""
    from idiaptts_b200 import ops
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(n + blocks)
    rows = 257
    x = rng.standard_normal((rows, blocks * n)).astype(np.float32)
    al = rng.uniform(-0.2, 0.2, rows).astype(np.float32)
    gy = rng.standard_normal((rows, blocks * n)).astype(np.float32)
    mean = rng.standard_normal(blocks * n).astype(np.float32)
    std = rng.uniform(0.5, 2.0, blocks * n).astype(np.float32)
    y_ref = glue_np.allpass_warp_forward(x.astype(np.float64), al.astype(np.float64), n)
    xd, ad, gd = torch.from_numpy(x).to(dev), torch.from_numpy(al).to(dev), torch.from_numpy(gy).to(dev)
    y = ops.allpass_forward(xd, ad, n).cpu().numpy()
    assert np.isfinite(y).all() and np.abs(y - y_ref).max() < 2e-5
    gx, ga = ops.allpass_backward(gd, xd, ad, n)
    gx_ref, ga_ref = glue_np.allpass_warp_backward(gy, x, al, n)
    assert np.abs(gx.cpu().numpy() - gx_ref).max() < 2e-5
    assert (np.abs(ga.cpu().numpy() - ga_ref) / (np.abs(ga_ref) + 1.0)).max() < 1e-4
    # with de-normalise / normalise folded in (AllPassWarpLayer.forward_fixed_alphas)
    md, sd = torch.from_numpy(mean).to(dev), torch.from_numpy(std).to(dev)
    yn = ops.allpass_forward(xd, ad, n, md, sd).cpu().numpy()
    yn_ref = (glue_np.allpass_warp_forward(x.astype(np.float64) * std + mean, al.astype(np.float64), n) - mean) / std
    assert np.abs(yn - yn_ref).max() < 5e-5
    gxn, gan = ops.allpass_backward(gd, xd, ad, n, md, sd)
    gxn_ref, gan_ref = glue_np.allpass_warp_backward(gy / std, x.astype(np.float64) * std + mean, al, n)
    assert np.abs(gxn.cpu().numpy() - gxn_ref * std).max() < 5e-5
    assert (np.abs(gan.cpu().numpy() - gan_ref) / (np.abs(gan_ref) + 1.0)).max() < 1e-4


def test_layer_module_interface():
    from idiaptts_b200.AllPassWarp import AllPassWarpLayer
    dev = torch.device("cuda", 0)
    n = 30
    cfg = AllPassWarpLayer.Config(alpha_layer_in_dims=[8], alpha_ranges=[0.2], batch_first=True, warp_matrix_size=n,
                                  mean=np.zeros(n, np.float32), std_dev=np.ones(n, np.float32))
    layer = cfg.create_model().to(dev)
    B, T = 4, 11
    feats = torch.randn(B, T, n, device=dev)
    emb = torch.randn(B, T, 8, device=dev)
    (out, combined, alpha), meta = layer((feats, emb), lengths=[T] * B, max_lengths=T)
    assert out.shape == feats.shape and combined.shape == (B, T, 1) and float(alpha.abs().max()) <= 0.2
    out.sum().backward()
    assert layer.alpha_layers[0].weight.grad is not None and torch.isfinite(layer.alpha_layers[0].weight.grad).all()
    assert all(not b.requires_grad for b in layer.buffers())
    y2, _ = layer.forward_sample(np.random.randn(T, n).astype(np.float32), np.full((T, 1), 0.1, np.float32))
    assert y2.shape == (1, T, n)


@pytest.mark.parametrize("n,blocks,norm", [(60, 1, False), (60, 1, True), (20, 1, True), (60, 3, False), (32, 1, False), (64, 1, True), (4, 1, False)])
def test_tensor_core_forward_vs_fp64_oracle(n, blocks, norm):
    """alpha constant over long runs of rows (one per speaker): tiles of 128 units that share alpha run on the tensor cores
    (3xTF32); a tile with ONE run boundary is processed as two single-alpha segments, a tile with more than two runs falls back to
    the recursion (the run lengths below produce all three cases); both against the fp64 oracle, and the two implementations against
    each other."""
    from idiaptts_b200 import ops
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(100 + n + blocks)
    runs = [300, 128, 5, 511, 77, 640]           # rows per speaker: boundaries fall inside and on tile edges
    rows = sum(runs)
    x = rng.standard_normal((rows, blocks * n)).astype(np.float32)
    al = np.concatenate([np.full(r, rng.uniform(-0.2, 0.2), np.float32) for r in runs])
    mean = rng.standard_normal(blocks * n).astype(np.float32) if norm else None
    std = rng.uniform(0.5, 2.0, blocks * n).astype(np.float32) if norm else None
    xd, ad = torch.from_numpy(x).to(dev), torch.from_numpy(al).to(dev)
    md = torch.from_numpy(mean).to(dev) if norm else None
    sd = torch.from_numpy(std).to(dev) if norm else None
    y_tc = ops.allpass_forward(xd, ad, n, md, sd, impl="tc").cpu().numpy()
    y_cc = ops.allpass_forward(xd, ad, n, md, sd, impl="cc").cpu().numpy()
    xin = x.astype(np.float64) * std + mean if norm else x.astype(np.float64)
    y_ref = glue_np.allpass_warp_forward(xin, al.astype(np.float64), n)
    if norm:
        y_ref = (y_ref - mean) / std
    assert np.isfinite(y_tc).all()
    assert np.abs(y_cc - y_ref).max() < 5e-5
    assert np.abs(y_tc - y_ref).max() < 5e-5
    assert np.abs(y_tc - y_cc).max() < 5e-5


@pytest.mark.parametrize("n,blocks,norm", [(60, 1, False), (60, 1, True), (20, 1, True), (60, 3, False), (64, 1, True), (4, 1, False)])   # n = 64: two-stage ring
def test_tensor_core_backward_vs_fp64_oracle(n, blocks, norm):
    """Backward of the same configurations: grad_x through the transposed matrix, grad_alpha through the tangent matrix (tiles
    with one run boundary as two segments, tiles with more runs through the recursion kernel)."""
    from idiaptts_b200 import ops
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(200 + n + blocks)
    runs = [300, 128, 5, 511, 77, 640]
    rows = sum(runs)
    x = rng.standard_normal((rows, blocks * n)).astype(np.float32)
    gy = rng.standard_normal((rows, blocks * n)).astype(np.float32)
    al = np.concatenate([np.full(r, rng.uniform(-0.2, 0.2), np.float32) for r in runs])
    mean = rng.standard_normal(blocks * n).astype(np.float32) if norm else None
    std = rng.uniform(0.5, 2.0, blocks * n).astype(np.float32) if norm else None
    xd, ad, gd = torch.from_numpy(x).to(dev), torch.from_numpy(al).to(dev), torch.from_numpy(gy).to(dev)
    md = torch.from_numpy(mean).to(dev) if norm else None
    sd = torch.from_numpy(std).to(dev) if norm else None
    gx_tc, ga_tc = ops.allpass_backward(gd, xd, ad, n, md, sd, impl="tc")
    gx_cc, ga_cc = ops.allpass_backward(gd, xd, ad, n, md, sd, impl="cc")
    if norm:
        gx_ref, ga_ref = glue_np.allpass_warp_backward(gy / std, x.astype(np.float64) * std + mean, al, n)
        gx_ref = gx_ref * std
    else:
        gx_ref, ga_ref = glue_np.allpass_warp_backward(gy, x, al, n)
    # grad_alpha is a 60-term dot product with large, cancelling terms (the tangent matrix grows with the cepstral index): the
    # fp32 recursion and the 3xTF32 path both reach 0.6 - 1.5e-4 of (|ref| + 1) on the worst row
    for (gx, ga), tol in (((gx_cc, ga_cc), 4e-4), ((gx_tc, ga_tc), 4e-4)):
        assert np.abs(gx.cpu().numpy() - gx_ref).max() < 5e-5
        assert (np.abs(ga.cpu().numpy() - ga_ref) / (np.abs(ga_ref) + 1.0)).max() < tol
