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
