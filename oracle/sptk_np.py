"""CPU oracle (TEST INFRASTRUCTURE, not product code): numpy restatement of the SPTK routines that
IdiapTTS reaches through pysptk.

pysptk (PyPI, un-pinned, reference requirements.txt:11) wraps r9y9/SPTK (mcep.c, freqt.c, frqtr in
mcep.c, mgc2sp.c, theq.c); neither is present under /root/reference, so the published algorithms are
restated in fp64 and anchored on the reference's call sites and goldens:

  * call sites  idiaptts/src/data_preparation/audio/AudioProcessing.py:146 (pysptk.mcep, etype=1 eps=1e-8
                itype=3), :252 (pysptk.mgc2sp gamma=0), :40 (pysptk.util.mcepalpha)
  * goldens     test/integration/fixtures/WORLD/cmp_mcep20/*.cmp columns 0..19 (PINNED to 2e-6 by
                tests/test_oracle_golden.py together with oracle/world_np.cheaptrick)
  * mc2sp / mgc2sp / sp2mc at order 59: pinned only by round trips (reference test asserts
    sum-sq-err < 100, test_WorldFeatLabelGen.py:823).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import numpy as np


def freqt(c1, order, alpha):
    """pysptk.freqt(ceps, order, alpha): all-pass frequency transform c1[0..m1] -> g[0..order]."""
    c1 = np.asarray(c1, np.float64)
    m1 = len(c1) - 1
    m2 = order
    a = float(alpha)
    b = 1.0 - a * a
    g = np.zeros(m2 + 1)
    d = np.zeros(m2 + 1)
    for i in range(-m1, 1):
        d[:] = g
        g[0] = c1[-i] + a * d[0]
        if m2 >= 1:
            g[1] = b * d[0] + a * d[1]
        for j in range(2, m2 + 1):
            g[j] = d[j - 1] + a * (d[j] - g[j - 1])
    return g


def frqtr(c1, order, alpha):
    """SPTK frqtr (static in mcep.c): transform of the autocorrelation-like sequence r[0..m1] -> [0..order]."""
    c1 = np.asarray(c1, np.float64)
    m1 = len(c1) - 1
    m2 = order
    a = float(alpha)
    g = np.zeros(m2 + 1)
    d = np.zeros(m2 + 1)
    for i in range(-m1, 1):
        d[:] = g
        g[0] = c1[-i]
        for j in range(1, m2 + 1):
            g[j] = d[j - 1] + a * (d[j] - g[j - 1])
    return g


def freqt_matrix(m1, m2, alpha):
    """A with freqt(c, m2, alpha) == A @ c for c of order m1; shape [m2+1, m1+1] (vectorised over unit inputs)."""
    a = float(alpha)
    b = 1.0 - a * a
    n_in = m1 + 1
    g = np.zeros((m2 + 1, n_in))
    eye = np.eye(n_in)
    for i in range(-m1, 1):
        d = g.copy()
        g[0] = eye[-i] + a * d[0]
        if m2 >= 1:
            g[1] = b * d[0] + a * d[1]
        for j in range(2, m2 + 1):
            g[j] = d[j - 1] + a * (d[j] - g[j - 1])
    return g


def frqtr_matrix(m1, m2, alpha):
    a = float(alpha)
    n_in = m1 + 1
    g = np.zeros((m2 + 1, n_in))
    eye = np.eye(n_in)
    for i in range(-m1, 1):
        d = g.copy()
        g[0] = eye[-i]
        for j in range(1, m2 + 1):
            g[j] = d[j - 1] + a * (d[j] - g[j - 1])
    return g


class McepError(RuntimeError):
    pass


_MATRIX_CACHE = {}


def _mcep_matrices(m, f2, alpha):
    """The three all-pass maps the Newton loop applies, as fp64 matrices (identical maps to the recursions above;
    test_oracle_golden.py checks matrix == recursion). Speeds the oracle up ~100x over python recursions."""
    key = (m, f2, float(alpha))
    if key not in _MATRIX_CACHE:
        _MATRIX_CACHE[key] = (freqt_matrix(f2, m, alpha), freqt_matrix(m, f2, -alpha), frqtr_matrix(f2, 2 * m, alpha))
    return _MATRIX_CACHE[key]


def mcep_frame(amp, order, alpha, miniter=2, maxiter=30, threshold=0.001, eps=1e-8):
    """One frame of pysptk.mcep(x, order, alpha, etype=1, eps, itype=3): amp[K] amplitude spectrum -> mc[order+1].

    Returns (mc, iterations, converged)."""
    amp = np.asarray(amp, np.float64)
    K = len(amp)
    flng = 2 * (K - 1)
    f2 = flng // 2
    m = order
    per_h = amp * amp + eps
    per = np.concatenate((per_h, per_h[f2 - 1:0:-1]))
    if np.any(per <= 0.0):
        raise McepError("zero(s) are found in periodogram, use eps option to floor")
    c = np.fft.ifft(np.log(per)).real
    c[0] /= 2.0
    c[f2] /= 2.0
    A, B, R = _mcep_matrices(m, f2, alpha)
    mc = A @ c[:f2 + 1]  # freqt(c, m, alpha)
    s = c[0]
    al = (-alpha) ** np.arange(m + 1)
    ii, kk = np.meshgrid(np.arange(m + 1), np.arange(m + 1), indexing="ij")
    converged = False
    it = 0
    for j in range(1, maxiter + 1):
        it = j
        cc = np.zeros(flng)
        cc[:f2 + 1] = B @ mc  # freqt(mc, f2, -alpha)
        C = np.fft.fft(cc).real
        r = np.fft.ifft(per / np.exp(C + C)).real
        rt = R @ r[:f2 + 1]  # frqtr(r, 2m, alpha)
        t = rt[0]
        if j >= miniter:
            if abs((t - s) / t) < threshold:
                converged = True
                break
            s = t
        M = rt[np.abs(ii - kk)] + rt[ii + kk]
        try:
            dlt = np.linalg.solve(M, rt[:m + 1] - al)
        except np.linalg.LinAlgError as e:
            raise McepError("failed to compute mcep; error occured in theq") from e
        mc = mc + dlt
    return mc, it, converged


def mcep(x, order=25, alpha=0.35, miniter=2, maxiter=30, threshold=0.001, etype=0, eps=0.0, min_det=1.0e-6,
         itype=0):
    """pysptk.mcep signature; only the reference's mode (itype=3 amplitude in, etype=1 eps floor) is restated."""
    if itype != 3 or etype not in (0, 1):
        raise NotImplementedError("oracle restates itype=3, etype in (0, 1) only (AudioProcessing.py:146)")
    x = np.asarray(x, np.float64)
    e = eps if etype == 1 else 0.0
    if x.ndim == 1:
        return mcep_frame(x, order, alpha, miniter, maxiter, threshold, e)[0]
    return np.stack([mcep_frame(f, order, alpha, miniter, maxiter, threshold, e)[0] for f in x])


def mgc2sp(mc, alpha=0.0, gamma=0.0, fftlen=256):
    """pysptk.mgc2sp for gamma == 0: complex log spectrum [.., fftlen/2+1] (reference takes exp(real), AudioProcessing.py:252-256)."""
    if gamma != 0.0:
        raise NotImplementedError("gamma != 0 is SURVEY 8(f) N3")
    mc = np.asarray(mc, np.float64)
    if mc.ndim == 1:
        c = freqt(mc, fftlen // 2, -alpha)
        return np.fft.rfft(c, fftlen)
    return np.stack([mgc2sp(f, alpha, gamma, fftlen) for f in mc])


def mc2sp(mc, alpha, fftlen):
    """pysptk.mc2sp: mel-cepstrum -> POWER spectrum."""
    mc = np.asarray(mc, np.float64)
    if mc.ndim == 1:
        c = freqt(mc, fftlen // 2, -alpha)
        c[0] *= 2.0
        sym = np.concatenate((c, c[fftlen // 2 - 1:0:-1]))
        return np.exp(np.fft.rfft(sym).real)
    return np.stack([mc2sp(f, alpha, fftlen) for f in mc])


def sp2mc(powerspec, order, alpha):
    """pysptk.sp2mc: power spectrum -> mel-cepstrum (linear, no Newton refinement)."""
    p = np.asarray(powerspec, np.float64)
    if p.ndim == 1:
        c = np.fft.irfft(np.log(p))
        c[0] /= 2.0
        return freqt(c, order, alpha)
    return np.stack([sp2mc(f, order, alpha) for f in p])


def mcepalpha(fs, start=0.0, stop=1.0, step=0.001, num_points=1000):
    """pysptk.util.mcepalpha: the all-pass constant whose warp best matches the mel scale at fs."""
    def _melscale_vector(fs, length):
        step_ = (fs / 2.0) / length
        mel = np.log(1 + step_ * np.arange(0, length) / 1000.0)
        return mel / mel[-1]

    def _warping_vector(alpha, length):
        step_ = np.pi / length
        omega = step_ * np.arange(0, length)
        num = (1 - alpha * alpha) * np.sin(omega)
        den = (1 + alpha * alpha) * np.cos(omega) - 2 * alpha
        warp = np.arctan(num / den)
        warp[warp < 0] += np.pi
        return warp / warp[-1]

    alpha_candidates = np.arange(start, stop, step)
    mel = _melscale_vector(fs, num_points)
    dist = [np.sqrt(np.sum((mel - _warping_vector(a, num_points)) ** 2) / num_points) for a in alpha_candidates]
    return alpha_candidates[int(np.argmin(dist))]
