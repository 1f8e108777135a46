"""CPU oracle (TEST INFRASTRUCTURE, not a product path) of the reference's MLPG: maximum-probability parameter generation
from [static | delta | delta-delta] means and a diagonal covariance, idiaptts/misc/mlpg.py:94-127 (MLPG.generation) with
build_win_mats (:31-52) and build_poe (:54-92).  The reference solves the banded normal equations with bandmat (not
installable here); this restatement builds the same pentadiagonal system and solves it with scipy.linalg.solveh_banded, and
`generation_dense` solves the dense system as an independent cross-check.  Parity is unpinned by reference goldens (the
reference tests hold no MLPG vectors); the two solvers agreeing to round-off is the anchor."""
import numpy as np
from scipy.linalg import solveh_banded

# windows of MLPG.generation (:95-99): (l, u, coefficients); row t of the window matrix holds the coefficients at t-l .. t+u
WINDOWS = ((0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5])), (1, 1, np.array([1.0, -2.0, 1.0])))
EDGE_VAR = 100000000000.0  # :113-116, the delta / delta-delta experts of the first and last frame are switched off


def _win_dense(frames):
    mats = []
    for l, u, c in WINDOWS:
        W = np.zeros((frames, frames))
        for t in range(frames):
            for o in range(-l, u + 1):
                if 0 <= t + o < frames:
                    W[t, t + o] = c[o + l]
        mats.append(W)
    return mats


def _frames_params(features, covariance, feature_dim, d):
    frames = features.shape[0]
    var = np.empty((frames, 3))
    mean = np.empty((frames, 3))
    for w in range(3):
        var[:, w] = covariance[feature_dim * w + d, feature_dim * w + d]
        mean[:, w] = features[:, feature_dim * w + d]
    var[0, 1] = var[0, 2] = var[-1, 1] = var[-1, 2] = EDGE_VAR
    return mean / var, 1.0 / var  # b_frames, tau_frames (:121-122)


def generation_dense(features, covariance, feature_dim):
    """P = sum_w W_w^T diag(tau_w) W_w, b = sum_w W_w^T b_w, x = P^-1 b with dense matrices."""
    features = np.asarray(features, np.float64)
    frames = features.shape[0]
    Ws = _win_dense(frames)
    out = np.zeros((frames, feature_dim))
    for d in range(feature_dim):
        bf, tf = _frames_params(features, covariance, feature_dim, d)
        P = sum(W.T @ (tf[:, w][:, None] * W) for w, W in enumerate(Ws))
        b = sum(W.T @ bf[:, w] for w, W in enumerate(Ws))
        out[:, d] = np.linalg.solve(P, b)
    return out


def generation(features, covariance, feature_dim):
    """Same system in banded form (bandwidth 2), solved by banded Cholesky like bandmat.linalg.solveh (:125)."""
    features = np.asarray(features, np.float64)
    covariance = np.asarray(covariance, np.float64)
    frames = features.shape[0]
    out = np.zeros((frames, feature_dim))
    for d in range(feature_dim):
        bf, tf = _frames_params(features, covariance, feature_dim, d)
        ab = np.zeros((3, frames))  # lower form: ab[k, i] = P[i + k, i]
        b = np.zeros(frames)
        for w, (l, u, c) in enumerate(WINDOWS):
            for t in range(frames):
                for oi in range(-l, u + 1):
                    i = t + oi
                    if not 0 <= i < frames:
                        continue
                    b[i] += c[oi + l] * bf[t, w]
                    for oj in range(oi, u + 1):
                        j = t + oj
                        if 0 <= j < frames:
                            ab[j - i, i] += tf[t, w] * c[oi + l] * c[oj + l]
        out[:, d] = solveh_banded(ab, b, lower=True)
    return out
