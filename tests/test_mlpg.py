"""MLPG: the oracle's banded system against a dense solve of the same normal equations (CPU), and the CUDA
pentadiagonal solver against the oracle (GPU)."""
import numpy as np
import pytest
import torch

from oracle import mlpg_np


def _case(rng, T, D):
    f = rng.standard_normal((T, 3 * D))
    cov = np.diag(rng.uniform(0.05, 3.0, 3 * D))
    return f, cov


@pytest.mark.parametrize("T,D", [(1, 2), (2, 1), (3, 2), (41, 3)])
def test_oracle_banded_equals_dense(T, D):
    f, cov = _case(np.random.default_rng(T * 10 + D), T, D)
    if T == 1:  # a single frame: only the static expert (and the switched-off edge experts) act
        x = mlpg_np.generation(f, cov, D)
        assert np.allclose(x, mlpg_np.generation_dense(f, cov, D), atol=1e-12)
        assert np.allclose(x[0], f[0, :D], atol=1e-9)
        return
    assert np.abs(mlpg_np.generation(f, cov, D) - mlpg_np.generation_dense(f, cov, D)).max() < 1e-10


def test_oracle_constant_trajectory_is_a_fixed_point():
    # static means constant, delta means zero: the smooth trajectory is the constant itself
    T, D = 30, 2
    f = np.zeros((T, 3 * D))
    f[:, :D] = [1.5, -0.7]
    x = mlpg_np.generation(f, np.eye(3 * D), D)
    assert np.abs(x - f[:, :D]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_mlpg_matches_oracle(dtype):
    from idiaptts_b200 import ops
    from idiaptts_b200.mlpg import MLPG
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    D = 7
    lens = [1, 2, 3, 57, 130, 4]
    cov = np.diag(rng.uniform(0.05, 3.0, 3 * D))
    feats = [rng.standard_normal((T, 3 * D)).astype(dtype) for T in lens]
    ref = np.concatenate([mlpg_np.generation(f.astype(np.float64), cov, D) for f in feats])
    off = torch.tensor(np.concatenate(([0], np.cumsum(lens))), dtype=torch.int64, device=dev)
    out = ops.mlpg(torch.from_numpy(np.concatenate(feats)).to(dev), torch.from_numpy(np.diag(cov).copy()).to(dev), off, D)
    assert out.dtype == torch.float64 and out.shape == (sum(lens), D)
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-9
    # the reference-facing class (one utterance, numpy in / out, full covariance matrix)
    x = MLPG().generation(feats[3], cov, D)
    assert x.dtype == np.float64 and np.abs(x - mlpg_np.generation(feats[3].astype(np.float64), cov, D)).max() < 1e-9


@pytest.mark.gpu
def test_gpu_mlpg_world_shaped_batch():
    """Acoustic-model-shaped output: 60 mel-cepstrum dimensions with deltas, 16 utterances; property check instead of the oracle
    at full size: the solution satisfies the normal equations P x = b (residual) for a sampled dimension."""
    from idiaptts_b200 import ops
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(4)
    D, U, T = 60, 16, 1301
    feats = rng.standard_normal((U * T, 3 * D)).astype(np.float32)
    var = rng.uniform(0.1, 2.0, 3 * D)
    off = torch.arange(U + 1, dtype=torch.int64, device=dev) * T
    out = ops.mlpg(torch.from_numpy(feats).to(dev), torch.from_numpy(var).to(dev), off, D).cpu().numpy()
    u, d = 5, 17
    ref = mlpg_np.generation(feats[u * T:(u + 1) * T].astype(np.float64)[:, [d, D + d, 2 * D + d]], np.diag(var[[d, D + d, 2 * D + d]]), 1)
    assert np.abs(out[u * T:(u + 1) * T, d] - ref[:, 0]).max() < 1e-9


@pytest.mark.gpu
def test_postprocess_world_applies_mlpg_per_feature():
    """WorldFeatLabelGen._postprocess_world (reference :357-415) on a network-output-shaped sample with deltas."""
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    rng = np.random.default_rng(9)
    D, nap, T = 20, 2, 83
    gen = WorldFeatLabelGen(add_deltas=True, num_coded_sps=D, num_bap=nap)
    gen.covs = [np.diag(rng.uniform(0.1, 2.0, 3 * D)), np.diag(rng.uniform(0.1, 2.0, 3)), np.atleast_1d(1.0),
                np.diag(rng.uniform(0.1, 2.0, 3 * nap))]
    sample = rng.standard_normal((T, 3 * D + 3 + 1 + 3 * nap))
    sample[:, 3 * D + 3] = rng.uniform(0, 1, T)
    out = gen._postprocess_world(sample.copy(), apply_mlpg=True)
    assert out.shape == (T, D + 1 + 1 + nap)
    assert np.abs(out[:, :D] - mlpg_np.generation(sample[:, :3 * D], gen.covs[0], D)).max() < 1e-9
    assert np.abs(out[:, D:D + 1] - mlpg_np.generation(sample[:, 3 * D:3 * D + 3], gen.covs[1], 1)).max() < 1e-9
    assert np.array_equal(out[:, D + 1], (sample[:, 3 * D + 3] > 0.5).astype(np.float64))
    assert np.abs(out[:, D + 2:] - mlpg_np.generation(sample[:, -3 * nap:], gen.covs[3], nap)).max() < 1e-9
    plain = gen._postprocess_world(sample.copy(), apply_mlpg=False)
    assert np.array_equal(plain[:, :D], sample[:, :D]) and np.array_equal(plain[:, D + 2:], sample[:, -3 * nap:][:, :nap])


@pytest.mark.gpu
def test_gpu_mlpg_shared_factor_table_edges():
    """The factor table shared by all utterances of a dimension (rows 0 .. T - 3 do not depend on T; the recurrence is cut off once
    its state repeats): utterance lengths on both sides of the in-place / tabulated switch (T < 6), lengths around the chunk size,
    and lengths far beyond the point where the factors become stationary, in one ragged batch, against the oracle."""
    from idiaptts_b200 import ops
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(8)
    D = 5
    lens = [5, 6, 7, 8, 9, 15, 16, 17, 1, 700, 2, 1301, 3, 64]
    var = rng.uniform(0.01, 5.0, 3 * D)
    var[0], var[D], var[2 * D] = 1e-4, 10.0, 10.0      # a sharply peaked static expert: factors stationary almost at once
    var[1], var[D + 1], var[2 * D + 1] = 10.0, 1e-3, 1e-3  # dominated by the dynamic experts: slow convergence
    feats = [rng.standard_normal((T, 3 * D)) for T in lens]
    ref = np.concatenate([mlpg_np.generation(f, np.diag(var), D) for f in feats])
    off = torch.tensor(np.concatenate(([0], np.cumsum(lens))), dtype=torch.int64, device=dev)
    x = torch.from_numpy(np.concatenate(feats)).to(dev)
    out = ops.mlpg(x, torch.from_numpy(var).to(dev), off, D).cpu().numpy()
    scale = np.abs(ref).max(axis=0)
    assert (np.abs(out - ref) / scale).max() < 1e-9
    # the same utterance alone and inside the batch: identical bits (the table does not depend on the batch composition)
    for u in (0, 3, 9):
        lo, hi = int(off[u]), int(off[u + 1])
        alone = ops.mlpg(x[lo:hi].contiguous(), torch.from_numpy(var).to(dev), torch.tensor([0, hi - lo], dtype=torch.int64, device=dev), D)
        assert np.array_equal(alone.cpu().numpy(), out[lo:hi])


def _ldl_rows(T, t0, t1, t2, rows):
    """The factor recurrence of csrc/mlpg.cu in numpy: (l1, l2, d) of the first `rows` rows of the pentadiagonal precision matrix of
    an utterance of T frames (T = None: no final frame in sight)."""
    te = 1.0 / 100000000000.0
    end = (lambda t: False) if T is None else (lambda t: t == T - 1)
    out_of = (lambda t: t < 0) if T is None else (lambda t: t < 0 or t >= T)
    tau1 = lambda t: 0.0 if out_of(t) else (te if (t == 0 or end(t)) else t1)  # noqa: E731
    tau2 = lambda t: 0.0 if out_of(t) else (te if (t == 0 or end(t)) else t2)  # noqa: E731
    d1 = d2 = l1p = 0.0
    res = []
    for i in range(rows):
        pii = t0 + 0.25 * (tau1(i - 1) + tau1(i + 1)) + tau2(i - 1) + 4.0 * tau2(i) + tau2(i + 1)
        pi1 = -2.0 * (tau2(i - 1) + tau2(i)) if i >= 1 else 0.0
        pi2 = tau2(i - 1) - 0.25 * tau1(i - 1) if i >= 2 else 0.0
        l2 = pi2 / d2 if i >= 2 else 0.0
        l1 = (pi1 - l2 * d2 * l1p) / d1 if i >= 1 else 0.0
        di = pii - l1 * l1 * d1 - l2 * l2 * d2
        res.append((l1, l2, di))
        d2, d1, l1p = d1, di, l1
    return res


@pytest.mark.parametrize("taus", [(1.0, 1.0, 1.0), (10.0, 0.3, 2.0), (0.1, 100.0, 100.0)])
def test_factor_rows_before_the_last_two_do_not_depend_on_the_length(taus):
    """The claim the CUDA solver's shared factor table rests on: rows 0 .. T - 3 of the L D L^T factorisation are the same for every
    utterance length (bit for bit), and once the recurrence state repeats it stays put; the factors agree with a dense LDL^T of the
    oracle's precision matrix."""
    t0, t1, t2 = taus
    free = _ldl_rows(None, t0, t1, t2, 400)
    for T in (6, 7, 9, 40, 400):
        own = _ldl_rows(T, t0, t1, t2, T)
        assert own[:T - 2] == free[:T - 2]                         # exact equality of Python floats (IEEE doubles)
        assert own[T - 2:] != free[T - 2:T]                        # ... and the last two rows really do differ
    # stationarity: the first row whose state equals the previous row's fixes all later rows
    fixed = next((i for i in range(4, 400) if free[i][2] == free[i - 1][2] == free[i - 2][2] and free[i][0] == free[i - 1][0]), None)
    if fixed is not None:
        assert all(r == free[fixed] for r in free[fixed:])
    # against the dense factorisation of the oracle's matrix (T = 12)
    T = 12
    cov = np.diag([1.0 / t0, 1.0 / t1, 1.0 / t2])
    _, tf = mlpg_np._frames_params(np.zeros((T, 3)), cov, 1, 0)
    P = sum(W.T @ (tf[:, w][:, None] * W) for w, W in enumerate(mlpg_np._win_dense(T)))
    L = np.linalg.cholesky(P)
    own = _ldl_rows(T, t0, t1, t2, T)
    np.testing.assert_allclose([r[2] for r in own], np.diag(L) ** 2, rtol=1e-9)                    # d_i
    np.testing.assert_allclose([r[0] for r in own[1:]], np.diag(L, -1) / np.diag(L)[:-1], rtol=1e-8, atol=1e-12)   # l1_i
    np.testing.assert_allclose([r[1] for r in own[2:]], np.diag(L, -2) / np.diag(L)[:-2], rtol=1e-8, atol=1e-12)   # l2_i
