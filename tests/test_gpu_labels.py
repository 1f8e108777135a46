"""lf0 / vuv preparation, deltas and statistics kernels: bit-exact against the oracle (which is pinned to the reference's
interpolate_lin / np.gradient) including the reference's edge cases."""
import numpy as np
import pytest
import torch

from oracle import glue_np

pytestmark = pytest.mark.gpu


def _tracks():
    rng = np.random.default_rng(0)
    tr = [np.zeros(7), np.full(5, 120.0), np.array([0, 0, 150.0, 0, 0, 180.0, 0]), np.array([0, 150.0, 0, 181.0]),
          np.array([140.0, 0, 0, 0]), np.array([0.0]), np.array([99.0]), np.array([0, 101.0]), np.array([100.0, 0, 130.0]),
          np.array([0, 0, 29.9, 30.0, 30.1, 0, 200.0, 0, 0]), np.array([25.0, 0, 0, 71.0, 0])]
    for _ in range(60):
        n = int(rng.integers(1, 80))
        tr.append(rng.uniform(60.0, 400.0, n) * (rng.uniform(size=n) > rng.uniform()))
    return tr


def test_lf0_vuv_bit_exact():
    from idiaptts_b200 import ops
    dev = torch.device("cuda", 0)
    tracks = _tracks()
    f0 = torch.from_numpy(np.concatenate(tracks)).to(dev)
    off = torch.from_numpy(np.concatenate(([0], np.cumsum([len(t) for t in tracks]))).astype(np.int64)).to(dev)
    lf0, vuv = ops.lf0_vuv(f0, off)
    lf0, vuv = lf0.cpu().numpy(), vuv.cpu().numpy()
    o = off.cpu().numpy()
    for u, tr in enumerate(tracks):
        ref_l, ref_v = glue_np.interpolate_lin(glue_np.lf0_from_f0(tr))
        assert np.array_equal(vuv[o[u]:o[u + 1]], ref_v.astype(np.float32)), tr  # vuv: bit-exact
        np.testing.assert_allclose(lf0[o[u]:o[u + 1]], ref_l, rtol=0, atol=1.5e-6, err_msg=str(tr))  # a few ulp of a float32 log (numpy's logf is not correctly rounded)
    # strided output (straight into packed feature rows)
    feats = torch.zeros((f0.numel(), 5), dtype=torch.float32, device=dev)
    flat = feats.view(-1)
    ops.lf0_vuv(f0, off, lf0_out=flat[1:], vuv_out=flat[3:], out_stride=5)
    assert np.array_equal(feats[:, 1].cpu().numpy(), lf0[:, 0]) and np.array_equal(feats[:, 3].cpu().numpy(), vuv[:, 0])
    assert float(feats[:, [0, 2, 4]].abs().sum()) == 0.0


def test_deltas_bit_exact_and_stats(golden):
    from idiaptts_b200 import ops
    dev = torch.device("cuda", 0)
    ids = ["LJ001-0002", "LJ001-0008", "LJ001-0006"]
    cm = [golden[i + "/cmp"] for i in ids]
    x = np.ascontiguousarray(np.concatenate([c[:, :20] for c in cm]))
    off = torch.from_numpy(np.concatenate(([0], np.cumsum([len(c) for c in cm]))).astype(np.int64)).to(dev)
    d, dd = ops.deltas(torch.from_numpy(x).to(dev), off)
    assert np.array_equal(d.cpu().numpy(), np.concatenate([c[:, 20:40] for c in cm]))   # reference-produced deltas
    assert np.array_equal(dd.cpu().numpy(), np.concatenate([c[:, 40:60] for c in cm]))  # and double deltas
    # two-frame utterance: one-sided differences at both ends
    two = torch.tensor([[1.0, 5.0], [4.0, 3.0]], dtype=torch.float32, device=dev)
    d2, dd2 = ops.deltas(two, torch.tensor([0, 2], dtype=torch.int64, device=dev))
    assert np.array_equal(d2.cpu().numpy(), np.gradient(two.cpu().numpy(), axis=0).astype(np.float32))
    assert np.array_equal(dd2.cpu().numpy(), np.zeros((2, 2), np.float32))
    # statistics: sum, sum of squares, Gram matrix in fp64
    full = torch.from_numpy(np.ascontiguousarray(np.concatenate([c[:, :60] for c in cm]))).to(dev)
    sums = torch.zeros(120, dtype=torch.float64, device=dev)
    gram = torch.zeros(3600, dtype=torch.float64, device=dev)
    ops.stats_accumulate(full, sums, gram)
    ops.stats_accumulate(full, sums, gram)  # accumulates
    f64 = full.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(sums.cpu().numpy()[:60], 2 * f64.sum(0), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(sums.cpu().numpy()[60:], 2 * (f64 ** 2).sum(0), rtol=1e-12)
    np.testing.assert_allclose(gram.cpu().numpy().reshape(60, 60), 2 * f64.T @ f64, rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize("dim", [64, 20, 4, 7, 1])
def test_deltas_ragged_edge_cases_match_numpy_gradient(dim):
    """Both delta kernels (float4 rows with a sliding window for dim % 4 == 0, element-wise otherwise) against
    utils.compute_deltas = np.gradient in float32 (idiaptts/misc/utils.py:103-105) on utterances of 1, 2, 3, 4, 5 frames, empty
    utterances, lengths around the rows-per-thread boundary and a long one; bit-exact."""
    from idiaptts_b200 import ops
    from oracle import glue_np
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(dim)
    lengths = [0, 1, 0, 0, 2, 3, 4, 5, 15, 16, 17, 0, 31, 32, 33, 1301, 1, 0]
    xs = [(rng.standard_normal((n, dim)) * 3).astype(np.float32) for n in lengths]
    x = torch.from_numpy(np.concatenate(xs)).to(dev)
    off = torch.from_numpy(np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)).to(dev)
    d, dd = ops.deltas(x, off)

    def grad(a):
        if len(a) < 2:
            return np.zeros_like(a)
        return glue_np.compute_deltas(a)
    ref_d = np.concatenate([grad(a) for a in xs])
    ref_dd = np.concatenate([grad(grad(a)) for a in xs])
    assert np.array_equal(d.cpu().numpy(), ref_d) and np.array_equal(dd.cpu().numpy(), ref_dd)
    d_only, none = ops.deltas(x, off, want_double=False)
    assert none is None and np.array_equal(d_only.cpu().numpy(), ref_d)
