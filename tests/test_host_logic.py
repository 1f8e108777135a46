"""Host-side logic that needs no GPU: layout conversion, normalisation extractors and their file formats, sharding,
the synthetic corpus generator, wav IO and the 'no CPU fallback' behaviour of the reference-facing entry points."""
import os

import numpy as np
import pytest
import torch

from idiaptts_b200 import distributed, synthetic
from idiaptts_b200.AudioProcessing import AudioProcessing
from idiaptts_b200.MeanCovarianceExtractor import MeanCovarianceExtractor
from idiaptts_b200.MeanStdDevExtractor import MeanStdDevExtractor
from idiaptts_b200.Synthesiser import Synthesiser
from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
from oracle import glue_np

IDS = ["LJ001-%04d" % i for i in range(1, 10)]


def test_convert_roundtrip_and_vuv_threshold():
    rng = np.random.default_rng(0)
    sample = rng.standard_normal((11, 63)).astype(np.float32)
    sample[:, 61] = rng.uniform(0, 1, 11)
    sample[3, 61] = 0.5
    a = WorldFeatLabelGen.convert_to_world_features(sample, num_coded_sps=60, num_bap=1)
    b = glue_np.convert_to_world_features(sample, num_coded_sps=60, num_bap=1)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    assert a[2][3] == 1.0  # vuv >= 0.5 -> 1 (WorldFeatLabelGen.py:754-756)
    back = WorldFeatLabelGen.convert_from_world_features(*a)
    assert back.shape == (11, 63) and np.array_equal(back[:, :61], sample[:, :61])
    # deltas are detected automatically
    wide = rng.standard_normal((5, 3 * 62 + 1)).astype(np.float32)
    c, l, v, bp = WorldFeatLabelGen.convert_to_world_features(wide, contains_deltas=False, num_coded_sps=60, num_bap=1)
    assert np.array_equal(l, wide[:, 180]) and np.array_equal(c, wide[:, :60])
    with pytest.raises(ValueError, match="WORLD requires all features"):
        WorldFeatLabelGen.convert_to_world_features(wide[:, :100], num_coded_sps=60, num_bap=1)


def test_trim_to_shortest():
    f = [np.zeros((10, 2)), np.zeros((8, 1)), None, np.zeros((9, 1))]
    out = WorldFeatLabelGen.trim_to_shortest(f)
    assert [None if o is None else len(o) for o in out] == [8, 8, None, 8]


def test_mean_std_extractor_matches_reference_bins(golden, tmp_path):
    ext = MeanStdDevExtractor()
    for id_ in IDS:
        ext.add_sample(golden[id_ + "/cmp"][:, :20])
    mean, std = ext.get_params()
    ref = golden["stats/mcep20/mean-std_dev/data"]
    np.testing.assert_allclose(mean, ref[0], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(std, ref[1], rtol=2e-5, atol=2e-6)
    # pre-reduced fp64 sums (the GPU path) give the same parameters
    ext2 = MeanStdDevExtractor()
    c = np.concatenate([golden[i + "/cmp"][:, :20] for i in IDS]).astype(np.float64)
    ext2.add_sums(len(c), c.sum(0), (c ** 2).sum(0))
    m2, s2 = ext2.get_params()
    np.testing.assert_allclose(m2, ref[0], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(s2, ref[1], rtol=2e-5, atol=2e-6)
    # file formats + combination of two subsets == the whole (MeanStdDevExtractor.combine_stats semantics)
    a, b = MeanStdDevExtractor(), MeanStdDevExtractor()
    for id_ in IDS[:4]:
        a.add_sample(golden[id_ + "/cmp"][:, :20].astype(np.float64))
    for id_ in IDS[4:]:
        b.add_sample(golden[id_ + "/cmp"][:, :20].astype(np.float64))
    a.save(str(tmp_path / "a"))
    b.save(str(tmp_path / "b"))
    assert sorted(np.load(str(tmp_path / "a-stats.npz")).files) == ["sum_frames", "sum_length", "sum_squared_frames"]
    assert sorted(np.load(str(tmp_path / "a-mean-std_dev.npz")).files) == ["mean", "std_dev", "sum_length"]
    mean_c, std_c = MeanStdDevExtractor.combine_mean_std([str(tmp_path / "a-stats.npz"), str(tmp_path / "b-stats.npz")],
                                                         dir_out=str(tmp_path), save_txt=False)
    np.testing.assert_allclose(mean_c[0], m2, rtol=1e-12)
    np.testing.assert_allclose(std_c[0], s2, rtol=1e-10)
    lm, ls = MeanStdDevExtractor.load(str(tmp_path / "mean-std_dev.npz"))
    np.testing.assert_allclose(lm[0], m2, rtol=1e-6)
    # legacy .bin reader
    with open(tmp_path / "legacy.bin", "wb") as f:
        f.write(np.int32(11579).tobytes())
        f.write(ref.astype(np.float64).tobytes())
    bm, bs = MeanStdDevExtractor.load(str(tmp_path / "legacy.bin"))
    np.testing.assert_allclose(bm[0], ref[0], rtol=1e-6)


def test_mean_covariance_extractor(tmp_path):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((500, 6))
    e = MeanCovarianceExtractor()
    e.add_sample(x[:200])
    e.add_sample(x[200:])
    mean, cov, std = e.get_params()
    np.testing.assert_allclose(mean[0], x.mean(0), atol=1e-12)
    np.testing.assert_allclose(cov, np.cov(x.T, bias=True), atol=1e-12)
    e.save(str(tmp_path / "set"))
    m, c, s = MeanCovarianceExtractor.load(str(tmp_path / "set-mean-covariance.npz"))
    np.testing.assert_allclose(c, cov, atol=1e-6)


def test_shard_utterances_balances_and_covers():
    rng = np.random.default_rng(2)
    lens = rng.integers(200, 2000, 1001)
    for world in (1, 2, 4, 8):
        shards = distributed.shard_utterances(lens, world)
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(len(lens)))
        loads = np.array([lens[s].sum() for s in shards])
        assert loads.max() - loads.min() <= lens.max()
        assert (loads.max() - loads.min()) / loads.mean() < 0.005


def test_synthetic_corpus_is_deterministic_and_world_shaped():
    w1, f1 = synthetic.make_corpus(3, 22050, seed=2, mean_dur=1.5)
    w2, f2 = synthetic.make_corpus(3, 22050, seed=2, mean_dur=1.5)
    for a, b, fa, fb in zip(w1, w2, f1, f2):
        assert torch.equal(a, b) and np.array_equal(fa, fb)
        assert a.dtype == torch.int16 and a.abs().max() > 8000
        assert len(fa) == int(1000.0 * a.numel() / 22050 / 5.0) + 1
        voiced = fa > 0
        assert 0.3 < voiced.mean() < 0.95 and fa[voiced].min() >= 71.0 and fa[0] == 0 and fa[-1] == 0
    w3, _ = synthetic.make_corpus(2, 22050, seed=2, mean_dur=1.5, first_utt=1)
    assert torch.equal(w3[0], w1[1])  # sharding by first_utt reproduces the same utterances
    wv, fv = synthetic.make_corpus(4, 16000, seed=3, mean_dur=1.2, std_dur=0.4)
    assert len({w.numel() for w in wv}) > 1


def test_wav_io_roundtrip(tmp_path):
    x = (np.sin(np.arange(1600) * 0.05) * 0.4).astype(np.float32)
    Synthesiser.write_wav(str(tmp_path / "a.wav"), x, 16000)
    data, fs = AudioProcessing.read_wav(str(tmp_path / "a.wav"))
    assert fs == 16000 and data.dtype == np.int16 and len(data) == 1600
    raw, fs = AudioProcessing.get_raw(str(tmp_path / "a.wav"), preemphasis=0.97)
    ref = data.astype(np.float64) / 32768.0
    np.testing.assert_array_equal(raw, glue_np.preemphasis(ref, 0.97))
    np.testing.assert_allclose(AudioProcessing.depreemphasis(raw, 0.97), ref, atol=1e-12)
    np.testing.assert_array_equal(AudioProcessing.depreemphasis(raw, 0.97), glue_np.depreemphasis(raw, 0.97))


def test_scalar_mappings_and_out_of_scope_errors():
    assert AudioProcessing.fs_to_frame_length(16000) == 1024 and AudioProcessing.fs_to_frame_length(48000) == 2048
    assert AudioProcessing.fs_to_num_bap(16000) == 1 and AudioProcessing.fs_to_num_bap(22050) == 2
    assert abs(AudioProcessing.fs_to_mgc_alpha(16000) - 0.41) < 1e-9 and abs(AudioProcessing.fs_to_mgc_alpha(22050) - 0.455) < 1e-9
    with pytest.raises(NotImplementedError):
        AudioProcessing.extract_mgc(None)
    with pytest.raises(NotImplementedError):
        AudioProcessing.decode_sp(np.zeros((2, 60)), "mgc", 16000)
    if not torch.cuda.is_available():  # no cached F0 -> DIO + StoneMask on the device: without a GPU this fails loudly
        with pytest.raises(RuntimeError):
            WorldFeatLabelGen.world_extract_features(np.zeros(1600), 16000, 5)
    with pytest.raises(NotImplementedError):
        WorldFeatLabelGen(sp_type="mgc")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a GPU")
def test_product_path_fails_loudly_without_gpu():
    from idiaptts_b200.compat import pyworld, pysptk
    with pytest.raises(RuntimeError, match="CUDA"):
        pyworld.cheaptrick(np.zeros(1600), np.zeros(21), np.arange(21) * 0.005, 16000)
    with pytest.raises(RuntimeError, match="CUDA"):
        pysptk.mcep(np.ones((3, 513)), 59, 0.58, etype=1, eps=1e-8, itype=3)
    with pytest.raises(RuntimeError, match="CUDA"):
        WorldFeatLabelGen.world_extract_features(np.zeros(1600), 16000, 5, f0=np.zeros(21))


def test_f0_stage_host_helpers():
    """Host-side pieces of the DIO / StoneMask path (no GPU): chunking, band geometry and workspace size through the C ABI."""
    from idiaptts_b200 import _lib, ops
    from oracle import dio_np
    off = np.array([0, 10, 30, 35, 100, 101])
    assert list(ops._utt_chunks(off, 30)) == [(0, 2), (2, 3), (3, 4), (4, 5)]   # a longer utterance forms its own chunk
    assert list(ops._utt_chunks(off, 1000)) == [(0, 5)]
    assert list(ops._utt_chunks(np.array([0]), 10)) == []
    lib = _lib.load()
    assert lib.b2w_dio_num_bands(71.0, 800.0, 2.0) == len(dio_np.dio_bands(16000)) == 7
    assert lib.b2w_dio_num_bands(800.0, 71.0, 2.0) < 0
    for fs in (16000, 22050, 48000):
        bands = dio_np.dio_bands(fs)
        taps = 2 * dio_np.mround(fs / 50.0) + 1 + sum(4 * dio_np.mround(fs / b / 2.0) for b in bands)
        assert ops.dio_fir_taps(fs) == taps
        small = lib.b2w_dio_workspace_bytes(100000, 4, 500, fs, 71.0, 800.0, 2.0)
        big = lib.b2w_dio_workspace_bytes(200000, 4, 500, fs, 71.0, 800.0, 2.0)
        assert 0 < small < big and (big - small) <= 100000 * (8 + 16 * 7 + 1)
    # WORLD sizes DIO's FFT so that the circular convolution never wraps: the premise of the direct (linear) FIR kernels
    for n, fs in ((48000, 16000), (143325, 22050), (480000, 48000)):
        h0 = dio_np.mround(fs / bands[0] / 2.0)
        assert dio_np.dio_fft_size(n, fs) >= (n + 1) + 2 * dio_np.mround(fs / 50.0) + 1 + 4 * h0


def test_lf0labelgen_reader_protocol(tmp_path):
    """Host half of LF0LabelGen (world/LF0LabelGen.py:63-210): raw float32 files, normalisation parameters, pre- / post-processing."""
    from idiaptts_b200.LF0LabelGen import LF0LabelGen
    from idiaptts_b200.MeanStdDevExtractor import MeanStdDevExtractor
    rng = np.random.default_rng(0)
    out = tmp_path / "labels"
    os.makedirs(str(out / "lf0"))
    os.makedirs(str(out / "vuv"))
    lf0 = (5.0 + 0.3 * rng.standard_normal((50, 1))).astype(np.float32)
    vuv = (rng.random((50, 1)) > 0.4).astype(np.float32)
    lf0[:, 0].tofile(str(out / "lf0" / "utt1.lf0"))
    vuv[:, 0].tofile(str(out / "vuv" / "utt1.vuv"))
    sample = LF0LabelGen.load_sample("utt1", str(out))
    assert sample.shape == (50, 2) and np.array_equal(sample[:, :1], lf0) and np.array_equal(sample[:, 1:], vuv)
    assert LF0LabelGen.load_lf0("utt1", str(out)).shape == (50, 1) and LF0LabelGen.load_vuv("utt1", str(out)).shape == (50, 1)
    ext = MeanStdDevExtractor()
    ext.add_sample(lf0)
    ext.save(str(out / "lf0" / "train"))
    gen = LF0LabelGen(str(out))
    assert gen.preprocess_sample(sample) is None           # no normalisation parameters yet (the reference logs an error)
    mean, std = gen.get_normalisation_params(str(out), "train")
    assert mean.shape == (1, 2) and mean[0, 1] == 0.0 and std[0, 1] == 1.0   # vuv: mean 0, std 1
    np.testing.assert_allclose(mean[0, 0], lf0.mean(), rtol=1e-6)
    norm = gen["utt1"]
    assert norm.dtype == np.float32 and abs(float(norm[:, 0].mean())) < 1e-4
    np.testing.assert_allclose(gen.postprocess_sample(norm), sample, atol=1e-5)
    soft = sample.copy()
    soft[:, 1] = np.clip(soft[:, 1] + 0.3 * rng.standard_normal(50), 0, 1)
    l, v = LF0LabelGen.convert_to_world_features(soft)
    assert set(np.unique(v)) <= {0.0, 1.0} and np.array_equal(l, soft[:, 0])
    assert LF0LabelGen.trim_end_sample(sample, 5).shape == (45, 2) and np.array_equal(LF0LabelGen.trim_end_sample(sample, 5, reverse=True), sample[5:])
    assert LF0LabelGen.trim_end_sample(sample, 0) is sample
    if not torch.cuda.is_available():                      # extraction itself needs the GPU: fails loudly without one
        with pytest.raises(RuntimeError):
            gen.gen_data(str(tmp_path), None, id_list=[])
