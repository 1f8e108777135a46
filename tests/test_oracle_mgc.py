"""The generalised mel-cepstrum oracle (oracle/mgc_np.py).  PARITY UNPINNED against SPTK (absent); what CAN be
pinned is pinned here: at gamma = 0 the Newton iteration is the mel-cepstral analysis of oracle/sptk_np.py (which reproduces the
reference's fixtures), the optimum is a stationary point of the published criterion, and the conversion recursions invert."""
import numpy as np
import pytest

from conftest import golden_utterance
from oracle import mgc_np, sptk_np, world_np


@pytest.fixture(scope="module")
def amp(golden):
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    sel = slice(40, 100)
    t = world_np.temporal_positions(len(f0))
    return np.sqrt(world_np.cheaptrick(x, f0[sel], t[sel], fs))


def test_gamma_zero_is_sptk_mcep_iterate_by_iterate(amp):
    for fr in (5, 30, 55):
        tr = []
        mgc_np.mgcep_frame(amp[fr], 19, 0.58, 0.0, eps=1e-8, miniter=7, maxiter=7, trace=tr)
        for steps in (1, 2, 3, 4):
            # mcep_frame applies `maxiter` Newton steps when it never converges (threshold 0)
            mc, _, _ = sptk_np.mcep_frame(amp[fr], 19, 0.58, eps=1e-8, miniter=steps, maxiter=steps, threshold=0.0)
            assert np.abs(mc - tr[steps][0]).max() < 1e-9, (fr, steps)
        full, _, _ = sptk_np.mcep_frame(amp[fr], 19, 0.58, eps=1e-8)              # SPTK's own stopping rule
        mine, _, _ = mgc_np.mgcep_frame(amp[fr], 19, 0.58, 0.0, eps=1e-8)
        assert np.abs(full - mine).max() < 2e-4                                    # stop one iterate apart at most


@pytest.mark.parametrize("gamma", [-1.0 / 3.0, -0.5, -0.2])
def test_optimum_is_a_stationary_point_of_the_criterion(amp, gamma):
    per = amp[30] ** 2 + 1e-8
    c, it, conv = mgc_np.mgcep_frame(amp[30], 24, 0.41, gamma, eps=1e-8, threshold=1e-13, maxiter=60)
    assert conv and it < 30
    E0 = mgc_np.uels(c, per, 0.41, gamma, 1024)
    rng = np.random.default_rng(0)
    for _ in range(20):   # a minimum: every small perturbation raises E (convex for -1 <= gamma <= 0)
        d = rng.standard_normal(25) * 1e-3
        assert mgc_np.uels(c + d, per, 0.41, gamma, 1024) > E0
    # finite-difference gradient ~ 0
    g = np.array([(mgc_np.uels(c + 1e-6 * e, per, 0.41, gamma, 1024) - mgc_np.uels(c - 1e-6 * e, per, 0.41, gamma, 1024)) / 2e-6
                  for e in np.eye(25)])
    assert np.abs(g).max() < 1e-6
    # SPTK-style stopping (threshold 1e-3) lands within 1e-3 of it
    c3, it3, _ = mgc_np.mgcep_frame(amp[30], 24, 0.41, gamma, eps=1e-8)
    assert it3 <= it
    if gamma > -1.0:   # (the all-pole case gamma = -1 has a nearly flat criterion on this frame: close in E, not in c)
        assert np.abs(c3 - c).max() < 2e-3
    assert abs(mgc_np.uels(c3, per, 0.41, gamma, 1024) - E0) < 2e-3


def test_conversions(amp):
    c, _, _ = mgc_np.mgcep_frame(amp[10], 59, 0.41, -1.0 / 3.0, eps=1e-8)
    # pysptk.mgc2sp route (mgc2mgc -> cepstrum -> FFT) == the definition of the model
    a_def = mgc_np.mgc_amplitude(c, 0.41, -1.0 / 3.0, 1024)
    a_sptk = np.exp(mgc_np.mgc2sp(c, 0.41, -1.0 / 3.0, 1024).real)
    assert np.abs(a_def / a_sptk - 1).max() < 1e-7
    # the model follows the envelope it was fitted to (order 59: ~1-2 dB rms)
    assert np.sqrt(np.mean((20 * np.log10(a_def / amp[10])) ** 2)) < 3.0
    # gnorm / ignorm and gc2gc there-and-back, mc2b / b2mc
    for g in (-1.0 / 3.0, 0.0):
        assert np.abs(mgc_np.ignorm(mgc_np.gnorm(c, g), g) - c).max() < 1e-12
    n = mgc_np.gnorm(c, -1.0 / 3.0)
    back = mgc_np.gc2gc(mgc_np.gc2gc(n, -1.0 / 3.0, 200, -0.5), -0.5, 59, -1.0 / 3.0)
    assert np.abs(back - n).max() < 1e-9
    assert np.abs(mgc_np.b2mc(mgc_np.mc2b(c, 0.41), 0.41) - c).max() < 1e-12
    # gamma = 0 entry point agrees with sptk_np.mgc2sp
    mc, _, _ = sptk_np.mcep_frame(amp[10], 59, 0.41, eps=1e-8)
    assert np.abs(mgc_np.mgc2sp(mc, 0.41, 0.0, 1024).real - sptk_np.mgc2sp(mc, 0.41, 0.0, 1024).real).max() < 1e-9


def test_merlin_post_filter_keeps_the_energy(amp):
    mc = np.stack([sptk_np.mcep_frame(amp[f], 59, 0.41, eps=1e-8)[0] for f in (10, 30)])
    pf = mgc_np.merlin_post_filter(mc, 0.41)
    assert np.array_equal(pf[:, 1], mc[:, 1]) and np.allclose(pf[:, 2:], 1.4 * mc[:, 2:], rtol=1e-12)
    for t in range(2):
        r0 = mgc_np.c2acr_r0(sptk_np.freqt(mc[t], 511, -0.41), 1024)
        r1 = mgc_np.c2acr_r0(sptk_np.freqt(pf[t], 511, -0.41), 1024)
        assert abs(r1 / r0 - 1) < 1e-9
