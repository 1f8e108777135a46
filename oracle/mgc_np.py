"""CPU oracle (TEST INFRASTRUCTURE, not product code) for the generalised mel-cepstrum branch of the reference
(sp_type = "mgc"):

    AudioProcessing.extract_mgc   -> pysptk.mgcep(amp_sp, order, alpha, gamma = -1/3, eps = 1e-8, etype = 1, itype = 3)
                                     (idiaptts/src/data_preparation/audio/AudioProcessing.py:123-140)
    AudioProcessing.mgc_to_amp_sp -> exp(Re pysptk.mgc2sp(mgc, alpha, gamma, fftlen))                       (:259-275)
    AudioProcessing.decode_sp(post_filtering=True) -> nnmnkwii.postfilters.merlin_post_filter               (:308-311)

PARITY UNPINNED, and more so than the rest of the oracle: pysptk / SPTK / nnmnkwii are absent from /root/reference and cannot
be installed; the reference holds no fixture for this branch (its own test skips it, test_WorldFeatLabelGen.py:417).  What is
restated here is the PUBLISHED definition, not SPTK's source text:

  * mel-generalised cepstral analysis (Tokuda, Kobayashi, Masuko, Imai, "Mel-generalized cepstral analysis", ICSLP 1994): the
    coefficients c(0..M) of  H(z) = (1 + gamma sum_m c(m) z~^-m)^(1/gamma),  z~^-1 = (z^-1 - alpha) / (1 - alpha z^-1),  that
    minimise the UELS criterion  E = mean_w [ I(w) / |H(w)|^2 + log |H(w)|^2 ]  (I = periodogram) -- a convex problem for
    -1 <= gamma <= 0.  `mgcep` runs an exact Newton iteration on E (gradient and Hessian in closed form, Hessian = Toeplitz +
    Hankel like SPTK's), starting from SPTK's initial value (linear mel-cepstrum converted with gc2gc), and stops like SPTK
    (relative change of epsilon = exp(E - 1) below `threshold`, at least `miniter`, at most `maxiter` iterations).  SPTK
    iterates on gain-normalised coefficients, so individual iterates differ; both converge to the same minimiser.
    Pin available here: at gamma = 0 the iteration IS the mel-cepstral analysis of oracle/sptk_np.py, which reproduces the
    reference's fixtures (tests/test_oracle_mgc.py checks iterate-by-iterate equality).
  * gnorm / ignorm / gc2gc / mgc2mgc / mc2b / b2mc / c2acr: the standard SPTK recursions (SPTK reference manual).
  * merlin_post_filter: restated from the nnmnkwii documentation (Merlin's post-filter: scale c(2..) by 1.4, keep the energy).
"""
import numpy as np

from . import sptk_np


# ---- warped-frequency tables -----------------------------------------------------------------------------------------
def warped_omega(fftlen, alpha):
    """w~(w) of the first-order all-pass at the bins w_j = 2 pi j / fftlen, j = 0 .. fftlen/2."""
    w = 2.0 * np.pi * np.arange(fftlen // 2 + 1) / fftlen
    return w + 2.0 * np.arctan2(alpha * np.sin(w), 1.0 - alpha * np.cos(w))


def _tables(order_max, fftlen, alpha):
    wt = warped_omega(fftlen, alpha)
    k = np.arange(order_max + 1)[:, None]
    return np.cos(k * wt[None, :]), np.sin(k * wt[None, :])


def bin_weights(fftlen):
    """mean over the full circle from the half spectrum: weights 1 for DC / Nyquist, 2 otherwise, / fftlen."""
    w = np.full(fftlen // 2 + 1, 2.0)
    w[0] = w[-1] = 1.0
    return w / fftlen


# ---- SPTK recursions -----------------------------------------------------------------------------------------------------
def gnorm(c, gamma):
    c = np.array(c, np.float64)
    if gamma != 0.0:
        k = 1.0 + gamma * c[0]
        c[1:] = c[1:] / k
        c[0] = k ** (1.0 / gamma)
    else:
        c[0] = np.exp(c[0])
    return c


def ignorm(c, gamma):
    c = np.array(c, np.float64)
    if gamma != 0.0:
        k = c[0] ** gamma
        c[1:] = c[1:] * k
        c[0] = (k - 1.0) / gamma
    else:
        c[0] = np.log(c[0])
    return c


def gc2gc(c1, g1, m2, g2):
    """generalised cepstrum (normalised form) of order len(c1)-1, gamma g1 -> order m2, gamma g2"""
    c1 = np.asarray(c1, np.float64)
    m1 = len(c1) - 1
    c2 = np.zeros(m2 + 1)
    c2[0] = c1[0]
    for i in range(1, m2 + 1):
        ss1 = ss2 = 0.0
        mn = m1 if m1 < i else i - 1
        for k in range(1, mn + 1):
            mk = i - k
            cc = c1[k] * c2[mk]
            ss2 += k * cc
            ss1 += mk * cc
        if i <= m1:
            c2[i] = c1[i] + (g2 * ss2 - g1 * ss1) / i
        else:
            c2[i] = (g2 * ss2 - g1 * ss1) / i
    return c2


def mgc2mgc(c1, a1, g1, m2, a2, g2):
    a = (a2 - a1) / (1.0 - a1 * a2)
    c1 = np.asarray(c1, np.float64)
    if a == 0.0:
        c = gnorm(c1, g1)
        c = gc2gc(c, g1, m2, g2)
        return ignorm(c, g2)
    c = sptk_np.freqt(c1, m2, a)
    c = gnorm(c, g1)
    c = gc2gc(c, g1, m2, g2)
    return ignorm(c, g2)


def mc2b(mc, alpha):
    mc = np.asarray(mc, np.float64)
    b = np.array(mc, np.float64)
    for m in range(len(mc) - 2, -1, -1):
        b[..., m] = mc[..., m] - alpha * b[..., m + 1]
    return b


def b2mc(b, alpha):
    b = np.asarray(b, np.float64)
    mc = np.array(b, np.float64)
    for m in range(len(b.T) - 2, -1, -1):
        mc[..., m] = b[..., m] + alpha * b[..., m + 1]
    return mc


def c2acr_r0(c, fftlen):
    """r(0) of pysptk.c2acr(c, 0, fftlen): the mean of exp(2 Re FFT(c)) over the circle."""
    c = np.asarray(c, np.float64)
    buf = np.zeros(fftlen)
    buf[:len(c)] = c
    return float(np.mean(np.exp(2.0 * np.fft.fft(buf).real)))


# ---- analysis -------------------------------------------------------------------------------------------------------------
class MgcepError(RuntimeError):
    pass


def mgc_amplitude(c, alpha, gamma, fftlen, tables=None):
    """|H(w_j)|, j = 0 .. fftlen/2, straight from the definition of the model."""
    c = np.asarray(c, np.float64)
    cs, sn = tables if tables is not None else _tables(len(c) - 1, fftlen, alpha)
    C = c @ cs[:len(c)] - 1j * (c @ sn[:len(c)])
    if gamma == 0.0:
        return np.exp(C.real)
    return np.abs(1.0 + gamma * C) ** (1.0 / gamma)


def uels(c, per, alpha, gamma, fftlen):
    """E(c) = mean_w [ I / |H|^2 + log |H|^2 ]"""
    amp = mgc_amplitude(c, alpha, gamma, fftlen)
    return float(np.sum(bin_weights(fftlen) * (per / amp ** 2 + 2.0 * np.log(amp))))


def mgcep_frame(amp, order, alpha, gamma, miniter=2, maxiter=30, threshold=0.001, eps=1e-8, trace=None):
    """One frame of pysptk.mgcep(x, order, alpha, gamma, etype=1, eps, itype=3): amp[K] -> mgc[order+1] (otype 0).
    Returns (mgc, iterations, converged)."""
    amp = np.asarray(amp, np.float64)
    K = len(amp)
    fftlen = 2 * (K - 1)
    m = order
    per = amp * amp + eps
    if np.any(per <= 0.0):
        raise MgcepError("zero(s) are found in periodogram, use eps option to floor")
    W = bin_weights(fftlen)
    cs, sn = _tables(2 * m, fftlen, alpha)
    # initial value (SPTK): linear mel-cepstrum of the log periodogram, converted to gamma
    full = np.concatenate((per, per[K - 2:0:-1]))
    c0 = np.fft.ifft(np.log(full)).real
    c0[0] /= 2.0
    c0[fftlen // 2] /= 2.0
    A0 = sptk_np.freqt_matrix(fftlen // 2, m, alpha)
    mc = A0 @ c0[:fftlen // 2 + 1]   # cepstrum of the log POWER with c(0), c(N/2) halved = coefficients of log |H| (as in mcep)
    c = mgc2mgc(mc, alpha, 0.0, m, alpha, gamma) if gamma != 0.0 else mc.copy()
    ii, kk = np.meshgrid(np.arange(m + 1), np.arange(m + 1), indexing="ij")
    prev, converged, it = None, False, 0
    for j in range(1, maxiter + 1):
        it = j
        Cre, Cim = c @ cs[:m + 1], -(c @ sn[:m + 1])
        if gamma == 0.0:
            Gre, Gim, g2 = np.ones(K), np.zeros(K), np.ones(K)
            P = per * np.exp(-2.0 * Cre)
            logH2 = 2.0 * Cre
        else:
            Gre, Gim = 1.0 + gamma * Cre, gamma * Cim
            g2 = Gre * Gre + Gim * Gim
            P = per * g2 ** (-1.0 / gamma)
            logH2 = np.log(g2) / gamma
        E = float(np.sum(W * (P + logH2)))
        # SPTK's epsilon = mean I / |D|^2 of the gain-normalised model D = H / K, K^2 = exp(mean log |H|^2) (H is minimum phase)
        epsilon = float(np.sum(W * P)) * np.exp(float(np.sum(W * logH2)))
        if trace is not None:
            trace.append((c.copy(), E))
        if j >= miniter and prev is not None:
            if abs((epsilon - prev) / epsilon) < threshold:
                converged = True
                break
        prev = epsilon
        qre, qim = Gre / g2, -Gim / g2                 # q = conj(G) / |G|^2 = 1 / G
        a = W * 2.0 * (1.0 - P)
        grad = cs[:m + 1] @ (a * qre) + sn[:m + 1] @ (a * qim)         # 2 (1 - P) Re(e^{-j m w~} / G)
        t = cs[:m + 1] @ (W * 2.0 * P / g2)                            # Toeplitz part: 2 P / |G|^2 cos(k w~)
        beta = W * (2.0 * P - 2.0 * gamma * (1.0 - P))
        q2re, q2im = qre * qre - qim * qim, 2.0 * qre * qim
        h = cs @ (beta * q2re) + sn @ (beta * q2im)                    # Hankel part: beta Re(e^{-j k w~} / G^2), k <= 2 m
        Hm = t[np.abs(ii - kk)] + h[ii + kk]
        try:
            step = np.linalg.solve(Hm, grad)
        except np.linalg.LinAlgError as e:
            raise MgcepError("failed to compute mgcep; singular Newton system") from e
        c = c - step
    return c, it, converged


def mgcep(x, order=25, alpha=0.35, gamma=0.0, miniter=2, maxiter=30, threshold=0.001, etype=0, eps=0.0, min_det=1.0e-6,
          itype=0, otype=0):
    """pysptk.mgcep signature; only the reference's mode (itype=3 amplitude in, etype=1 eps floor, otype=0) is restated."""
    if itype != 3 or etype not in (0, 1) or otype != 0:
        raise NotImplementedError("oracle restates itype=3, etype in (0, 1), otype=0 only (AudioProcessing.py:131-138)")
    x = np.asarray(x, np.float64)
    e = eps if etype == 1 else 0.0
    if x.ndim == 1:
        return mgcep_frame(x, order, alpha, gamma, miniter, maxiter, threshold, e)[0]
    return np.stack([mgcep_frame(f, order, alpha, gamma, miniter, maxiter, threshold, e)[0] for f in x])


# ---- synthesis side ----------------------------------------------------------------------------------------------------
def mgc2sp(mgc, alpha=0.0, gamma=0.0, fftlen=256):
    """pysptk.mgc2sp: complex log spectrum [.., fftlen/2+1] = FFT of mgc2mgc(mgc -> alpha 0, gamma 0, order fftlen/2)
    (the reference takes exp(real), AudioProcessing.py:268-275)."""
    mgc = np.asarray(mgc, np.float64)
    if mgc.ndim == 1:
        c = mgc2mgc(mgc, alpha, gamma, fftlen // 2, 0.0, 0.0)
        return np.fft.rfft(c, fftlen)
    return np.stack([mgc2sp(f, alpha, gamma, fftlen) for f in mgc])


def mgc_to_amp_sp(mgc, fs, alpha, gamma=-1.0 / 3.0, n_fft=1024):
    """AudioProcessing.mgc_to_amp_sp (:259-275)"""
    return np.exp(mgc2sp(np.ascontiguousarray(mgc, np.float64), alpha, gamma, n_fft).real.astype(np.float32))


def merlin_post_filter(mgc, alpha, minimum_phase_order=511, fftlen=1024, coef=1.4, weight=None):
    """nnmnkwii.postfilters.merlin_post_filter: scale the mel-cepstrum (from c(2) on) by `coef`, then restore the energy r(0) of
    the minimum-phase impulse response through b(0)."""
    mgc = np.asarray(mgc, np.float64)
    T, D = mgc.shape
    if weight is None:
        weight = np.ones(D) * coef
        weight[:2] = 1.0
    out = np.empty_like(mgc)
    for t in range(T):
        r0 = c2acr_r0(sptk_np.freqt(mgc[t], minimum_phase_order, -alpha), fftlen)
        p_r0 = c2acr_r0(sptk_np.freqt(mgc[t] * weight, minimum_phase_order, -alpha), fftlen)
        b = mc2b(mgc[t] * weight, alpha)
        b[0] = np.log(r0 / p_r0) / 2.0 + b[0]
        out[t] = b2mc(b, alpha)
    return out
