"""Pins the C oracle (the CPU baseline) against the reference's golden vectors and against the numpy oracle."""
import numpy as np
import pytest

from conftest import golden_utterance
from oracle import glue_np, sptk_np, world_c, world_np


@pytest.mark.parametrize("id_", ["LJ001-0002", "LJ001-0008", "LJ001-0004"])
def test_c_oracle_reproduces_reference_cmp(golden, id_):
    c = golden[id_ + "/cmp"]
    f0 = np.where(c[:, 63] > 0, np.exp(c[:, 60].astype(np.float64)), 0.0)
    feats = world_c.extract(golden[id_ + "/wav"], 16000, f0, 20, 0.58, preemphasis=0.97)
    assert feats.shape == (c.shape[0], 23)
    assert np.abs(feats[:, :20] - c[:, :20]).max() < 2e-6
    assert np.abs(feats[:, 22] - c[:, 64]).max() < 3e-5
    assert np.array_equal(feats[:, 21], c[:, 63])


def test_c_oracle_equals_numpy_oracle(golden):
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    t = world_np.temporal_positions(len(f0))
    sp_np = world_np.cheaptrick(x, f0, t, fs)
    sp_c = world_c.cheaptrick(x, f0, t, fs)
    assert (np.abs(sp_c - sp_np) / sp_np).max() < 1e-9
    v_np, co_np = world_np.d4c_coarse(x, f0, t, fs)
    v_c, co_c = world_c.d4c_coarse(x, f0, t, fs)
    assert np.array_equal(v_np, v_c) and np.abs(co_np[v_np] - co_c[v_np]).max() < 1e-8
    bap_np = world_np.code_aperiodicity(world_np.d4c(x, f0, t, fs), fs)
    np.testing.assert_allclose(world_c.bap_from_coarse(co_c, v_c, fs, 1024), bap_np, atol=1e-8)
    ref = [sptk_np.mcep_frame(np.sqrt(a), 59, 0.41, eps=1e-8) for a in sp_np]
    mc_c, it_c = world_c.mcep(np.sqrt(sp_np), 59, 0.41)
    assert np.abs(mc_c - np.stack([r[0] for r in ref])).max() < 1e-9
    assert np.array_equal(it_c, np.array([r[1] for r in ref]))
    lf0_np, vuv_np = glue_np.interpolate_lin(glue_np.lf0_from_f0(f0))
    lf0_c, vuv_c = world_c.lf0_vuv(f0)
    assert np.array_equal(vuv_c, vuv_np[:, 0].astype(np.float32))
    np.testing.assert_allclose(lf0_c, lf0_np[:, 0], atol=1.5e-6)


def test_c_synthesis_equals_numpy_oracle(golden):
    """The C port of the synthesis half (what bench.py's CPU arm times) against oracle/world_np.py::synthesize, both through
    pyworld.synthesize's signature and through the feature-level entry (Synthesiser.run_world_synth's per-utterance body)."""
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    T = 160
    x, f0 = x[:int(T * 0.005 * fs)], f0[40:40 + T].copy()
    t = world_np.temporal_positions(T)
    sp = world_np.cheaptrick(x, f0, t, fs)
    ap = world_np.d4c(x, f0, t, fs)
    y_np = world_np.synthesize(f0, sp, ap, fs)
    y_c = world_c.synthesize(f0, sp, ap, fs)
    assert y_c.shape == y_np.shape
    snr = 10 * np.log10((y_np ** 2).sum() / max(((y_c - y_np) ** 2).sum(), 1e-300))
    assert snr > 200, snr
    np.testing.assert_allclose(world_c.decode_aperiodicity(world_np.code_aperiodicity(ap, fs), fs, 1024),
                               world_np.decode_aperiodicity(world_np.code_aperiodicity(ap, fs), fs, 1024), rtol=1e-13)
    # feature level: [mcep60 | lf0 | vuv | bap] rows -> waveform
    mc = sptk_np.mcep(np.sqrt(sp), order=59, alpha=0.58, eps=1e-8, etype=1, itype=3).astype(np.float32)
    lf0, vuv = glue_np.interpolate_lin(glue_np.lf0_from_f0(f0))
    bap = world_np.code_aperiodicity(ap, fs).astype(np.float32)
    feats = np.concatenate((mc, lf0.astype(np.float32), vuv.astype(np.float32), bap), axis=1)
    y_ref = glue_np.world_features_to_raw(glue_np.mcep_to_amp_sp(mc, fs, alpha=0.58), lf0[:, 0].copy(), vuv[:, 0].copy(), bap.copy(), fs)
    y_feat = world_c.synthesize_features(feats, fs, 60, 0.58)
    assert y_feat.shape == y_ref.shape
    snr = 10 * np.log10((y_ref.astype(np.float64) ** 2).sum() / max(((y_feat.astype(np.float64) - y_ref) ** 2).sum(), 1e-300))
    assert snr > 120, snr   # float32 output samples
