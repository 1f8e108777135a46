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
    np.testing.assert_allclose(np.squeeze(mean_c), m2, rtol=1e-12)
    np.testing.assert_allclose(np.squeeze(std_c), s2, rtol=1e-10)
    lm, ls = MeanStdDevExtractor.load(str(tmp_path / "mean-std_dev.npz"))
    np.testing.assert_allclose(np.squeeze(lm), m2, rtol=1e-6)
    # legacy .bin reader
    with open(tmp_path / "legacy.bin", "wb") as f:
        f.write(np.int32(11579).tobytes())
        f.write(ref.astype(np.float64).tobytes())
    bm, bs = MeanStdDevExtractor.load(str(tmp_path / "legacy.bin"))
    np.testing.assert_allclose(np.squeeze(bm), ref[0], rtol=1e-6)


def test_mean_covariance_extractor(tmp_path):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((500, 6))
    e = MeanCovarianceExtractor()
    e.add_sample(x[:200])
    e.add_sample(x[200:])
    mean, cov = e.get_params()                                   # two values, like the reference
    np.testing.assert_allclose(mean[0], x.mean(0), atol=1e-12)
    np.testing.assert_allclose(cov, np.cov(x.T, bias=True), atol=1e-12)
    np.testing.assert_allclose(e.get_std_dev(), x.std(0), atol=1e-12)
    e.save(str(tmp_path / "set"))
    assert sorted(np.load(str(tmp_path / "set-stats.npz")).files) == ["sum_frames", "sum_length", "sum_product_frames"]
    assert sorted(np.load(str(tmp_path / "set-mean-covariance.npz")).files) == ["covariance", "mean", "sum_length"]
    m, c, s = MeanCovarianceExtractor.load(str(tmp_path / "set-mean-covariance.npz"))
    assert m.shape == (6,) and c.shape == (6, 6) and s.shape == (6,) and m.dtype == np.float32   # squeezed like the reference
    np.testing.assert_allclose(c, cov, atol=1e-6)
    np.testing.assert_allclose(s, x.std(0), atol=1e-6)
    # legacy .bin: int32 N, int32 rows, then [rows, d] float64 = mean row + covariance
    with open(tmp_path / "legacy-mean-covariance.bin", "wb") as f:
        f.write(np.array([500, 7], np.int32).tobytes())
        f.write(np.concatenate((mean, cov), axis=0).astype(np.float64).tobytes())
    bm, bc, bs = MeanCovarianceExtractor.load(str(tmp_path / "legacy-mean-covariance.bin"))
    np.testing.assert_allclose(bm, x.mean(0), atol=1e-6)
    np.testing.assert_allclose(bc, cov, atol=1e-6)
    # merging two subsets through their stats files, with a file-name prefix
    a, b = MeanCovarianceExtractor(), MeanCovarianceExtractor()
    a.add_sample(x[:123])
    b.add_sample(x[123:])
    a.save(str(tmp_path / "a"))
    b.save(str(tmp_path / "b"))
    mc, cc = MeanCovarianceExtractor.combine_mean_covariance([str(tmp_path / "a-stats.npz"), str(tmp_path / "b-stats.npz")],
                                                             dir_out=str(tmp_path), file_name="all", save_txt=False)
    np.testing.assert_allclose(cc, cov, atol=1e-12)
    assert (tmp_path / "all-mean-covariance.npz").exists() and (tmp_path / "all-stats.npz").exists()


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
    assert abs(AudioProcessing.mgc_gamma + 1.0 / 3.0) < 1e-15
    with pytest.raises(NotImplementedError):
        AudioProcessing.decode_sp(np.zeros((2, 60)), "mfbanks", 16000)     # other vocoder paths stay out of scope
    if not torch.cuda.is_available():  # the mgc branch is built and, like everything else, needs the GPU
        with pytest.raises(RuntimeError, match="CUDA"):
            AudioProcessing.extract_mgc(np.ones((3, 513)), fs=16000)
        with pytest.raises(RuntimeError, match="CUDA"):
            AudioProcessing.decode_sp(np.zeros((2, 60)), "mgc", 16000)
    if not torch.cuda.is_available():  # no cached F0 -> DIO + StoneMask on the device: without a GPU this fails loudly
        with pytest.raises(RuntimeError):
            WorldFeatLabelGen.world_extract_features(np.zeros(1600), 16000, 5)
    assert WorldFeatLabelGen(sp_type="mgc").dir_coded_sps == "mgc60"
    with pytest.raises(NotImplementedError):
        WorldFeatLabelGen(sp_type="mfbanks")


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


def _write_world_label_dir(root, rng, ids, D=6, nap=1, add_deltas=False, list_name="train"):
    """What WorldFeatLabelGen.gen_data leaves on disk (save_output :1121-1172 + the per-feature normalisation files), written
    with numpy only so the reader protocol can be tested without a GPU."""
    from idiaptts_b200.MeanCovarianceExtractor import MeanCovarianceExtractor as MC
    feats = {"mcep": ("mcep%d" % D, D), "lf0": ("lf0", 1), "vuv": ("vuv", 1), "bap": ("bap", nap)}
    data = {}
    for key, (sub, d) in feats.items():
        os.makedirs(os.path.join(root, sub), exist_ok=True)
        ext = MC() if add_deltas else MeanStdDevExtractor()
        for i in ids:
            T = 30 + 3 * len(i)
            x = (rng.standard_normal((T, d)) * (1.0 + np.arange(d)) + 0.5 * np.arange(d)).astype(np.float32)
            if key == "vuv":
                x = (x > 0).astype(np.float32)
                np.savez(os.path.join(root, sub, i), vuv=x)
                data[(key, i)] = x
                continue
            if add_deltas:
                dl, ddl = np.gradient(x, axis=0).astype(np.float32), np.gradient(np.gradient(x, axis=0), axis=0).astype(np.float32)
                np.savez(os.path.join(root, sub, i), **{key: x, key + "_deltas": dl, key + "_double_deltas": ddl})
                full = np.concatenate((x, dl, ddl), axis=1)
            else:
                np.savez(os.path.join(root, sub, i), **{key: x})
                full = x
            data[(key, i)] = full
            ext.add_sample(full)
        if key != "vuv":
            ext.save(os.path.join(root, sub, list_name + ("-deltas" if add_deltas else "")))
    return data


@pytest.mark.parametrize("add_deltas", [False, True])
def test_trainer_call_sequence_on_a_fresh_reader(tmp_path, add_deltas):
    """The sequence every reference trainer runs (AcousticModelTrainer.py:413-423, :439, :492): construct a reader on a label
    directory, get_normalisation_params(dir_out, file_name), __getitem__, postprocess_sample -- on a FRESH WorldFeatLabelGen (one
    that never ran gen_data), in the legacy keyword form and through WorldFeatLabelGen.Config.create_reader()."""
    rng = np.random.default_rng(3)
    ids = ["a", "bb", "ccc"]
    D, nap = 6, 1
    root = str(tmp_path / "labels")
    data = _write_world_label_dir(root, rng, ids, D, nap, add_deltas)
    reader = WorldFeatLabelGen(root, add_deltas=add_deltas, num_coded_sps=D, num_bap=nap)
    with pytest.raises(RuntimeError, match="get_normalisation_params"):
        reader["a"]                                           # no silent un-normalised data
    mean, std = reader.get_normalisation_params(root, "train")
    f3 = 3 if add_deltas else 1
    W = (D + 1 + nap) * f3 + 1
    assert np.asarray(mean).reshape(1, -1).shape == (1, W) and np.asarray(std).reshape(1, -1).shape == (1, W)
    mean, std = np.asarray(mean).reshape(-1), np.asarray(std).reshape(-1)
    vuv_col = (D + 1) * f3
    assert mean[vuv_col] == 0.0 and std[vuv_col] == 1.0      # vuv is never normalised (:616-618)
    raw = np.concatenate([data[(k, "bb")] for k in ("mcep", "lf0", "vuv", "bap")], axis=1)
    assert np.array_equal(reader.load("bb"), raw)
    x = reader["bb"]
    np.testing.assert_allclose(x, (raw - mean) / std, rtol=1e-5, atol=1e-6)
    # means really are the corpus means
    allrows = np.concatenate([np.concatenate([data[(k, i)] for k in ("mcep", "lf0", "vuv", "bap")], axis=1) for i in ids])
    keep = np.arange(W) != vuv_col
    np.testing.assert_allclose(mean[keep], allrows.mean(0)[keep], rtol=1e-4, atol=1e-5)
    if add_deltas:
        assert reader.covs[0].shape == (3 * D, 3 * D) and reader.covs[1].shape == (3, 3) and reader.covs[3].shape == (3 * nap, 3 * nap)
        assert reader.covs[2] is None
        back = reader.postprocess_sample(x.copy(), apply_mlpg=False)          # MLPG itself needs the GPU (tests/test_mlpg.py)
        assert back.shape == (len(raw), D + 2 + nap)
        np.testing.assert_allclose(back[:, :D], raw[:, :D], rtol=1e-4, atol=1e-4)
    else:
        back = reader.postprocess_sample(x.copy())
        np.testing.assert_allclose(back, raw, rtol=1e-4, atol=1e-4)
    # the Config form: fields, create_reader() loads the parameters, __getitem__ returns a dict keyed by the output name
    cfg = WorldFeatLabelGen.Config(name="acoustic_features", directory=root, norm_params_path=root, add_deltas=add_deltas,
                                   num_coded_sps=D, num_bap=nap)
    assert cfg.dir_labels == root and cfg.sp_type == "mcep" and cfg.load_bap
    r2 = WorldFeatLabelGen(cfg)
    assert r2.legacy_getitem is False and r2.norm_params is None
    r2.get_normalisation_params(root, "train")
    out = r2["bb"]
    assert set(out) == {"acoustic_features"}
    np.testing.assert_allclose(out["acoustic_features"], x, rtol=1e-6)
    # unknown keyword arguments are an error, the package's extensions are accepted
    with pytest.raises(TypeError):
        WorldFeatLabelGen(root, not_an_option=1)
    WorldFeatLabelGen(root, f0_cache={}, mgc_alpha=0.5)


def test_bind_host_to_gpu_is_harmless_without_nvml_or_gpu(monkeypatch):
    """distributed.bind_host_to_gpu is an optimisation only: without a GPU / NVML, or when switched off, it changes nothing."""
    from idiaptts_b200 import distributed
    before = os.sched_getaffinity(0)
    monkeypatch.setenv("B2W_NUMA_BIND", "0")
    assert distributed.bind_host_to_gpu(0) is None
    monkeypatch.setenv("B2W_NUMA_BIND", "1")
    got = distributed.bind_host_to_gpu(0)
    assert got is None or got <= before
    if got is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
