"""Pins the F0 oracle (DIO + StoneMask, oracle/dio_np.py) against the reference's golden vectors: columns 60 (lf0) and 63
(vuv) of test/integration/fixtures/WORLD/cmp_mcep20/*.cmp were produced by pyworld.wav2world (WorldFeatLabelGen.py:792-802)
from database/wav/*.wav with pre-emphasis 0.97."""
import numpy as np
import pytest

from conftest import golden_utterance
from oracle import dio_np, glue_np

IDS = ["LJ001-%04d" % i for i in range(1, 10)]


@pytest.mark.parametrize("id_", IDS)
def test_dio_stonemask_reproduce_reference_f0(golden, id_):
    x, c, _, fs = golden_utterance(golden, id_)
    f0, t = dio_np.wav2world_f0(x, fs)
    assert len(f0) == c.shape[0]
    assert np.array_equal(f0 > 0, c[:, 63] > 0)  # vuv: bit-exact
    v = f0 > 0
    assert np.abs(np.log(f0[v]).astype(np.float32) - c[v, 60]).max() < 1e-6  # float32 round-off of the stored lf0
    # the whole lf0 / vuv label columns (interpolate_lin included, WorldFeatLabelGen.py:798-802)
    lf0, vuv = glue_np.interpolate_lin(glue_np.lf0_from_f0(f0))
    assert np.array_equal(vuv[:, 0], c[:, 63])
    assert np.abs(lf0[:, 0] - c[:, 60]).max() < 2e-6


def test_dio_helpers():
    assert dio_np.suitable_fft_size(1000) == 1024 and dio_np.suitable_fft_size(1024) == 2048
    assert len(dio_np.dio_bands(16000)) == 7
    f = dio_np.low_cut_filter(641, 4096)
    assert abs(f.sum()) < 1e-12 and np.allclose(f[1:321], f[:-321:-1])  # unit DC rejection, symmetric about sample 0
    # short input: FixF0Contour leaves everything unvoiced
    f0, t = dio_np.dio(np.zeros(400), 16000)
    assert len(f0) == 6 and not f0.any()
    # pure tone: DIO + StoneMask find it
    fs = 16000
    n = np.arange(fs)
    x = 0.3 * np.sin(2 * np.pi * 150.0 * n / fs) + 0.1 * np.sin(2 * np.pi * 300.0 * n / fs)
    f0, t = dio_np.wav2world_f0(x, fs)
    mid = f0[40:160]
    assert np.all(mid > 0) and np.abs(mid - 150.0).max() < 0.5


def test_oracle_dio_is_homogeneous():
    """DIO only looks at zero crossings: scaling the waveform by a power of two (exact in IEEE arithmetic) changes nothing."""
    fs = 16000
    rng = np.random.default_rng(2)
    n = np.arange(fs)
    x = 0.3 * np.sin(2 * np.pi * (120.0 + 20.0 * n / fs) * n / fs) + 0.01 * rng.standard_normal(fs)
    a, t = dio_np.dio(x, fs)
    b, _ = dio_np.dio(0.25 * x, fs)
    assert np.array_equal(a, b) and (a > 0).sum() > 50
    # StoneMask adds 1e-12 to an amplitude sum (FixF0): homogeneous to ~1e-12 only
    np.testing.assert_allclose(dio_np.stonemask(x, a, t, fs), dio_np.stonemask(0.25 * x, b, t, fs), rtol=1e-9)
