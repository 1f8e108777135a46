"""Pins the CPU oracle against the reference's own golden vectors.

The reference's fixtures test/integration/fixtures/WORLD/cmp_mcep20/*.cmp were produced by the reference pipeline
(pyworld.wav2world -> pysptk.mcep / code_aperiodicity -> np.gradient deltas); the oracle must reproduce them to
float32 round-off before anything is compared with it."""
import os

import numpy as np
import pytest

from conftest import golden_utterance
from oracle import glue_np, sptk_np, world_np

IDS = ["LJ001-%04d" % i for i in range(1, 10)]


@pytest.mark.parametrize("id_", IDS)
def test_analysis_reproduces_reference_cmp(golden, id_):
    x, c, f0, fs = golden_utterance(golden, id_)
    T = c.shape[0]
    assert world_np.num_frames(len(x), fs) == T  # frame count: bit-exact
    t = world_np.temporal_positions(T)
    sp = world_np.cheaptrick(x, f0, t, fs)
    mc = sptk_np.mcep(np.sqrt(sp), order=19, alpha=0.58, eps=1e-8, etype=1, itype=3).astype(np.float32)
    assert np.abs(mc - c[:, :20]).max() < 2e-6
    assert glue_np.mcd_db(c[:, :20], mc) < 1e-5
    bap = world_np.code_aperiodicity(world_np.d4c(x, f0, t, fs), fs).astype(np.float32)
    assert np.abs(bap[:, 0] - c[:, 64]).max() < 3e-5
    # unvoiced constant (known-answer): 20*log10(1 - 1e-12)
    unv = c[:, 63] == 0
    assert np.all(bap[unv, 0] == np.float32(-8.685697e-12))


def test_deltas_reproduce_reference_cmp(golden):
    for id_ in IDS:
        c = golden[id_ + "/cmp"]
        assert np.array_equal(glue_np.compute_deltas(c[:, :20]), c[:, 20:40])
        assert np.array_equal(glue_np.compute_deltas(c[:, 20:40]), c[:, 40:60])
        assert np.array_equal(glue_np.compute_deltas(c[:, 64:65]), c[:, 65:66])


def test_stats_reproduce_reference_bins(golden):
    for feat, cols in (("mcep20", slice(0, 20)), ("lf0", slice(60, 61)), ("bap", slice(64, 65))):
        ext = glue_np.MeanStdDev()
        for id_ in IDS:
            if feat == "lf0":  # the lf0 stats fixture belongs to the WORLD/lf0 files (another F0 run), not to cmp col 60
                ext.add_sample(golden[id_ + "/lf0_other"][:, None])
            else:
                ext.add_sample(golden[id_ + "/cmp"][:, cols])
        n = int(golden["stats/%s/mean-std_dev/n" % feat])
        assert ext.sum_length == n == 11579
        ref = golden["stats/%s/mean-std_dev/data" % feat]
        mean, std = ext.get_params()
        np.testing.assert_allclose(mean, ref[0], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(std, ref[1], rtol=2e-5, atol=2e-6)
        raw = golden["stats/%s/stats/data" % feat]
        np.testing.assert_allclose(ext.sum_frames, raw[0], rtol=2e-5, atol=1e-2)
        np.testing.assert_allclose(ext.sum_squared_frames, raw[1], rtol=2e-5, atol=1e-2)


def test_interpolate_lin_known_answers(golden):
    """WORLD/lf0 + WORLD/vuv are interpolate_lin outputs for another F0 run: masking lf0 by vuv and re-running
    the interpolation must reproduce lf0 (idempotence on reference data)."""
    for id_ in IDS:
        lf0 = golden[id_ + "/lf0_other"]
        vuv = golden[id_ + "/vuv_other"]
        masked = np.where(vuv > 0, lf0, np.float32(0.0)).astype(np.float32)
        out, v = glue_np.interpolate_lin(masked)
        assert np.array_equal(v[:, 0].astype(np.float32), vuv)
        np.testing.assert_allclose(out[:, 0], lf0, rtol=0, atol=2e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/idiaptts"), reason="reference tree not present")
def test_interpolate_lin_matches_reference_function():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/idiaptts/misc/utils.py")
    ref = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(ref)
    except Exception as e:  # pragma: no cover
        pytest.skip("reference utils not importable: %r" % (e,))
    rng = np.random.default_rng(0)
    cases = [np.zeros(7, np.float32), np.full(5, 4.5, np.float32), np.array([0, 0, 5, 0, 0, 6, 0], np.float32),
             np.array([0, 5, 0, 6], np.float32), np.array([5, 0, 0, 0], np.float32), np.array([0], np.float32),
             np.array([4], np.float32), np.array([0, 4], np.float32), np.array([4, 0, 5], np.float32)]
    for _ in range(40):
        n = int(rng.integers(1, 60))
        v = (rng.uniform(3.5, 6.0, n) * (rng.uniform(size=n) > rng.uniform())).astype(np.float32)
        cases.append(v)
    for v in cases:
        a, av = ref.interpolate_lin(v)
        b, bv = glue_np.interpolate_lin(v)
        assert np.array_equal(np.asarray(a, np.float32), np.asarray(b, np.float32)), v
        assert np.array_equal(av, bv)


def test_matrix_forms_equal_recursions():
    rng = np.random.default_rng(1)
    c = rng.standard_normal(64)
    np.testing.assert_allclose(sptk_np.freqt_matrix(63, 19, 0.42) @ c, sptk_np.freqt(c, 19, 0.42), rtol=0, atol=1e-12)
    np.testing.assert_allclose(sptk_np.frqtr_matrix(63, 38, 0.42) @ c, sptk_np.frqtr(c, 38, 0.42), rtol=0, atol=1e-12)
    mc = rng.standard_normal(20) * 0.1
    np.testing.assert_allclose(sptk_np.freqt_matrix(19, 63, -0.42) @ mc, sptk_np.freqt(mc, 63, -0.42), rtol=0, atol=1e-12)


def test_constants():
    assert world_np.get_cheaptrick_fft_size(16000) == 1024
    assert world_np.get_cheaptrick_fft_size(22050) == 1024
    assert world_np.get_cheaptrick_fft_size(48000) == 2048
    assert [world_np.get_num_aperiodicities(fs) for fs in (16000, 22050, 48000)] == [1, 2, 5]
    assert world_np.get_d4c_fft_size(16000) == 2048 and world_np.get_d4c_fft_size(22050) == 2048
    assert abs(sptk_np.mcepalpha(16000) - 0.41) < 1e-9 and abs(sptk_np.mcepalpha(22050) - 0.455) < 1e-9
    np.testing.assert_allclose(world_np.xorshift_randn_sequence(4), [-1.32764, -0.622855, -1.609181, 1.179765], atol=1e-6)


def test_mc2sp_roundtrip_reference_threshold(golden):
    """Reference asserts sum((world_amp - mc2sp(mcep80))^2) < 100 (test_WorldFeatLabelGen.py:823)."""
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    T = c.shape[0]
    sp = world_np.cheaptrick(x, f0, world_np.temporal_positions(T), fs)
    amp = np.sqrt(sp)
    mc = glue_np.extract_mcep(amp, 80, 0.41)
    rec = glue_np.mcep_to_amp_sp(mc, fs, alpha=0.41)
    assert ((amp - rec) ** 2).sum() < 100
