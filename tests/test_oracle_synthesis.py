"""Properties of the oracle's synthesis half (PARITY UNPINNED by reference goldens): the checks the
reference itself makes (length, loose reconstruction error) plus internal consistency."""
import numpy as np

from conftest import golden_utterance
from oracle import glue_np, world_np


def test_reference_reconstruction_threshold(golden):
    """test_WorldFeatLabelGen.py:704-763: after peak normalisation sum((orig - WORLD resynthesis)^2) < 10000."""
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    T = len(f0)
    t = world_np.temporal_positions(T)
    sp = world_np.cheaptrick(x, f0, t, fs)
    ap = world_np.d4c(x, f0, t, fs)
    y = world_np.synthesize(f0, sp, ap, fs)
    assert len(y) == int(T * 5.0 * fs / 1000)
    assert abs(len(y) / fs / 0.005 - T) < 10  # test_AcousticModelTrainer.py:162-168
    n = min(len(x), len(y))
    a, b = x[:n] / np.abs(x[:n]).max(), y[:n] / np.abs(y[:n]).max()
    assert ((a - b) ** 2).sum() < 10000
    # energy is preserved within a few dB in voiced speech
    assert abs(10 * np.log10((y ** 2).sum() / (x[:n] ** 2).sum())) < 3.0


def test_pulse_structure(golden):
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    idx, shift, vuv = world_np.synthesis_time_base(f0, fs, 0.005, int(len(f0) * 5.0 * fs / 1000), 1024)
    assert np.all(np.diff(idx) > 0) and np.all((shift >= 0) & (shift <= 1.0 / fs + 1e-12))
    # unvoiced stretches tick at 500 Hz: 32 samples at 16 kHz
    d = np.diff(idx)
    assert (d == 32).mean() > 0.2


def test_codec_roundtrip_and_unvoiced_rule():
    fs, n = 22050, 1024
    bap = np.array([[-5.0, -2.0], [-0.3, -0.4], [-30.0, -10.0]])
    ap = world_np.decode_aperiodicity(bap, fs, n)
    assert np.all(ap[1] == 1.0 - 1e-12)  # mean(bap) > -0.5 -> unvoiced row
    back = world_np.code_aperiodicity(ap, fs)
    np.testing.assert_allclose(back[[0, 2]], bap[[0, 2]], atol=0.1)  # interpolation across the 3 kHz knot is not exact at 22.05 kHz
    ap16 = world_np.decode_aperiodicity(bap[:, :1], 16000, n)
    np.testing.assert_allclose(world_np.code_aperiodicity(ap16, 16000)[[0, 2]], bap[[0, 2], :1], atol=1e-9)


def test_depreemphasis_inverts_preemphasis():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1000)
    np.testing.assert_allclose(glue_np.depreemphasis(glue_np.preemphasis(x, 0.97), 0.97), x, atol=1e-9)


def test_exact_fmod_two_pi_argument():
    """The synthesis kernels compute fmod(x, 2 pi) as fma(-q, y, x) with q = floor(fl(x / y)) (q - 1 when the result is negative),
    csrc/synth.cu fmod_two_pi.  This is the C fmod bit for bit: the quotient is never under-estimated, and the remainder is a
    multiple of 2^-50 below 8, i.e. exactly representable, so the fused multiply-add does not round.  Checked here with exact
    rational arithmetic (the fma) against numpy's fmod, including arguments within an ulp of a multiple of 2 pi and the running
    phase WORLD produces for the 500 Hz unvoiced default at 16 kHz."""
    from fractions import Fraction
    rng = np.random.default_rng(0)
    y = 2.0 * 3.1415926535897932384
    xs = np.concatenate([rng.uniform(0, 2.5e5, 4000),
                         y * rng.integers(1, 30000, 4000) * (1 + rng.choice([-1, 0, 1], 4000) * 2.0 ** -52),
                         np.cumsum(np.full(3000, 2 * np.pi * 500 / 16000))])
    for x in xs:
        x = float(x)
        if x < y:
            r = x
        else:
            q = np.floor(x / y)
            exact = Fraction(x) - Fraction(q) * Fraction(y)
            r = float(exact)
            assert Fraction(r) == exact            # representable: the fma returns it unrounded
            if r < 0:
                exact = Fraction(x) - Fraction(q - 1) * Fraction(y)
                r = float(exact)
                assert Fraction(r) == exact
        assert r == float(np.fmod(x, y)) and 0.0 <= r < y


def test_exact_parallel_phase_scan_prototype():
    """scripts/exact_phase_scan_prototype.py: the running phase total[i] = fl(total[i-1] + inc[i]) evaluated by composing
    two-state integer maps (associative, hence parallel) is bit-identical to the sequential float64 loop -- also when many
    increments are exact half-ulp ties (round half to even depends on the parity of the running total) and across binade
    crossings.  Design aid for the next round's synthesis time-base kernel."""
    import importlib.util
    import math
    import os
    spec = importlib.util.spec_from_file_location(
        "exact_phase_scan_prototype", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts",
                                                   "exact_phase_scan_prototype.py"))
    proto = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(proto)
    rng = np.random.default_rng(5)
    # WORLD-like increments: voiced stretches with varying F0 between 500 Hz unvoiced defaults, three sampling rates
    for fs in (16000, 22050, 48000):
        inc = np.full(12000, 2.0 * math.pi * 500.0 / fs)
        inc[2000:6000] = 2.0 * math.pi * rng.uniform(71, 400, 4000) / fs
        assert np.array_equal(proto.sequential(inc), proto.exact_scan(inc, block=128))
    # forced ties: increments of the form (k + 1/2) ulp of the binade the total lives in, mixed with ordinary ones
    u = 2.0 ** (10 - 52)                      # ulp for totals in [2^10, 2^11)
    inc = np.concatenate(([1024.0 + 3 * u], (rng.integers(1, 1000, 3000) + 0.5) * u, rng.uniform(0, 1e-3, 3000),
                          (rng.integers(1, 2 ** 30, 3000) + 0.5) * u, rng.uniform(0.1, 1.0, 4000),
                          (rng.integers(1, 1000, 2000) + 0.5) * 2 * u))
    rng.shuffle(inc[1:])
    a, b = proto.sequential(inc), proto.exact_scan(inc, block=64)
    assert np.array_equal(a, b)
    assert a[-1] > 2048.0                     # the run crossed into the next binade (where half of the forced ties tie again)
