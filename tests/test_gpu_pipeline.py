"""End-to-end through the reference-facing entry points on the GPU: BASELINE config 1 (single synthetic 3 s 16 kHz utterance:
extract mcep60 alpha 0.58 + lf0/vuv/bap + WORLD resynthesis) against the oracle; gen_data with files; run_world_synth; and
size-independent properties at BASELINE-sized batches."""
import os

import numpy as np
import pytest
import torch

from oracle import glue_np, sptk_np, world_np

pytestmark = pytest.mark.gpu


def snr_db(ref, out):
    return 10 * np.log10((ref ** 2).sum() / max(((out - ref) ** 2).sum(), 1e-300))


def _write_corpus(tmp_path, n, fs, dur, seed):
    from idiaptts_b200 import synthetic
    from idiaptts_b200.Synthesiser import Synthesiser
    waves, f0s = synthetic.make_corpus(n, fs, seed=seed, mean_dur=dur, std_dur=0.2)
    ids, cache = [], {}
    os.makedirs(tmp_path / "wav", exist_ok=True)
    for u, (w, f) in enumerate(zip(waves, f0s)):
        id_ = "utt%03d" % u
        import wave
        with wave.open(str(tmp_path / "wav" / (id_ + ".wav")), "wb") as wf:
            wf.setnchannels(1)
            wf.setsampwidth(2)
            wf.setframerate(fs)
            wf.writeframes(w.numpy().tobytes())
        ids.append(id_)
        cache[id_] = f
    return ids, cache, waves, f0s


def test_config1_extract_and_resynthesis(tmp_path):
    """BASELINE.json configs[0]."""
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    from idiaptts_b200.AudioProcessing import AudioProcessing
    fs = 16000
    ids, cache, waves, f0s = _write_corpus(tmp_path, 1, fs, 3.0, seed=1)
    f0 = f0s[0]
    coded_sp, lf0, vuv, bap = WorldFeatLabelGen.extract_features(str(tmp_path / "wav"), ids[0], num_coded_sps=60, f0=f0, mgc_alpha=0.58)
    x = waves[0].numpy().astype(np.float64) / 32768.0
    T = world_np.num_frames(len(x), fs)
    assert coded_sp.shape == (T, 60) and lf0.shape == (T, 1) and vuv.shape == (T, 1) and bap.shape == (T, 1)
    assert coded_sp.dtype == lf0.dtype == vuv.dtype == bap.dtype == np.float32
    amp_ref, lf0_ref, vuv_ref, bap_ref = glue_np.world_extract_features(x, fs, 5, f0)
    mc_ref = glue_np.extract_mcep(amp_ref, 60, 0.58)
    assert np.array_equal(vuv, vuv_ref)                                   # vuv: bit-exact
    assert np.abs(lf0 - lf0_ref).max() < 1e-6
    assert glue_np.mcd_db(mc_ref, coded_sp) < 1e-3                        # tolerance 0.01 dB
    assert np.abs(bap - bap_ref).max() < 1e-4
    amp, _, _, _ = WorldFeatLabelGen.world_extract_features(x, fs, 5, f0=f0)
    assert amp.dtype == np.float64 and (np.abs(amp - amp_ref) / amp_ref).max() < 1e-6    # envelope tolerance 1e-4
    # resynthesis from the coded features
    amp_dec = AudioProcessing.decode_sp(coded_sp, "mcep", fs, alpha=0.58).astype(np.double)
    y = WorldFeatLabelGen.world_features_to_raw(amp_dec, lf0[:, 0].copy(), vuv[:, 0].copy(), bap, fs)
    y_ref = glue_np.world_features_to_raw(glue_np.mcep_to_amp_sp(coded_sp, fs, 0.58).astype(np.double), lf0[:, 0].copy(), vuv[:, 0].copy(), bap, fs)
    assert len(y) == len(y_ref) and snr_db(y_ref, y) > 60


@pytest.mark.parametrize("add_deltas", [False, True])
def test_gen_data_files_and_statistics(tmp_path, add_deltas):
    from idiaptts_b200.MeanStdDevExtractor import MeanStdDevExtractor
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    fs = 22050
    ids, cache, waves, f0s = _write_corpus(tmp_path, 4, fs, 0.7, seed=6)
    gen = WorldFeatLabelGen(str(tmp_path / "out"), add_deltas=add_deltas, num_coded_sps=60, num_bap=2, f0_cache=cache)
    label_dict, means, stds = gen.gen_data(str(tmp_path / "wav"), str(tmp_path / "out"), file_id_list="train.txt", id_list=ids,
                                           return_dict=True)
    width = 3 * (60 + 1 + 2) + 1 if add_deltas else 64
    for u, id_ in enumerate(ids):
        T = world_np.num_frames(len(waves[u]), fs)
        assert label_dict[id_].shape == (T, width)
        for d, key in (("mcep60", "mcep"), ("lf0", "lf0"), ("vuv", "vuv"), ("bap", "bap")):
            arch = np.load(str(tmp_path / "out" / d / (id_ + ".npz")))
            assert key in arch.files and arch[key].shape[0] == T
            if add_deltas and key != "vuv":
                assert key + "_deltas" in arch.files and key + "_double_deltas" in arch.files
                assert np.array_equal(arch[key + "_deltas"], glue_np.compute_deltas(arch[key]))
                assert np.array_equal(arch[key + "_double_deltas"], glue_np.compute_deltas(arch[key + "_deltas"]))
    # one utterance against the oracle
    x = waves[1].numpy().astype(np.float64) / 32768.0
    amp_ref, lf0_ref, vuv_ref, bap_ref = glue_np.world_extract_features(x, fs, 5, f0s[1])
    mc_ref = glue_np.extract_mcep(amp_ref, 60, sptk_np.mcepalpha(fs))
    mc = np.load(str(tmp_path / "out" / "mcep60" / (ids[1] + ".npz")))["mcep"]
    assert glue_np.mcd_db(mc_ref, mc) < 1e-3
    assert np.array_equal(np.load(str(tmp_path / "out" / "vuv" / (ids[1] + ".npz")))["vuv"], vuv_ref)
    # statistics == reference extractor fed with the saved features
    if not add_deltas:
        assert means.shape == (64,) and stds.shape == (64,) and means[61] == 0.0 and stds[61] == 1.0
        ext = MeanStdDevExtractor()
        for id_ in ids:
            ext.add_sample(np.load(str(tmp_path / "out" / "mcep60" / (id_ + ".npz")))["mcep"].astype(np.float64))
        m, s = ext.get_params()
        np.testing.assert_allclose(means[:60], m, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(stds[:60], s, rtol=1e-5, atol=1e-7)
        saved_m, saved_s = MeanStdDevExtractor.load(str(tmp_path / "out" / "mcep60" / "train-mean-std_dev.npz"))
        np.testing.assert_allclose(saved_m, m, rtol=1e-5, atol=1e-6)
        st = np.load(str(tmp_path / "out" / "lf0" / "train-stats.npz"))
        assert int(st["sum_length"]) == sum(world_np.num_frames(len(w), fs) for w in waves)
    else:
        assert means[0].shape == (1, 180) and stds[0].shape == (180, 180)
    # reader half: load_sample returns what gen_data wrote, __getitem__ normalises, postprocess_sample inverts (+ MLPG with deltas)
    loaded = WorldFeatLabelGen.load_sample(ids[2], str(tmp_path / "out"), add_deltas=add_deltas, num_coded_sps=60, num_bap=2)
    assert loaded.dtype == np.float32 and np.array_equal(loaded, label_dict[ids[2]])
    norm = gen[ids[2]]
    assert norm.shape == loaded.shape and np.isfinite(norm).all()
    back = gen.postprocess_sample(norm, apply_mlpg=True)
    assert back.shape == (loaded.shape[0], 64)
    static = loaded[:, np.r_[0:60, 180:181, 183:184, 184:186]] if add_deltas else loaded
    assert np.array_equal(back[:, 61], static[:, 61])                       # vuv survives exactly
    if add_deltas:
        # deltas computed from the statics are (nearly) consistent with them, so the maximum-probability trajectory stays close
        assert np.abs(back[:, :60] - static[:, :60]).max() < 0.5 and np.abs(back[:, 60] - static[:, 60]).max() < 0.5
        full = np.concatenate([label_dict[i][:, :180] for i in ids]).astype(np.float64)
        np.testing.assert_allclose(means[0][0], full.mean(0), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(stds[0], np.cov(full.T, bias=True), rtol=1e-5, atol=1e-6)
        assert os.path.exists(str(tmp_path / "out" / "mcep60" / "train-deltas-mean-covariance.npz"))
    else:
        np.testing.assert_allclose(back, static, rtol=1e-4, atol=1e-4)
    # a FRESH reader (never ran gen_data) finds the same parameters on disk: the trainer call sequence
    fresh = WorldFeatLabelGen(str(tmp_path / "out"), add_deltas=add_deltas, num_coded_sps=60, num_bap=2)
    fresh.get_normalisation_params(str(tmp_path / "out"), "train")
    np.testing.assert_allclose(fresh[ids[2]], norm, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(fresh.postprocess_sample(fresh[ids[2]], apply_mlpg=True), back, rtol=1e-4, atol=1e-4)
    # without an F0 cache the F0 stage of pyworld.wav2world (DIO + StoneMask) runs on the device
    ld2, _, _ = WorldFeatLabelGen(str(tmp_path / "o2"), num_coded_sps=60, num_bap=2).gen_data(str(tmp_path / "wav"), None, id_list=ids,
                                                                                            return_dict=True)
    assert [ld2[i].shape[0] for i in ids] == [label_dict[i].shape[0] for i in ids]


def test_run_world_synth_writes_reference_named_files(tmp_path):
    from types import SimpleNamespace
    from idiaptts_b200 import ops, pipeline, synthetic
    from idiaptts_b200.AudioProcessing import AudioProcessing
    from idiaptts_b200.Synthesiser import Synthesiser
    dev = torch.device("cuda", 0)
    fs = 16000
    waves, f0s = synthetic.make_corpus(2, fs, seed=8, mean_dur=0.6)
    an = pipeline.WorldAnalyzer(fs, 60, device=dev)
    feats, _, _ = an.extract(ops.RaggedBatch.from_host([w.numpy() for w in waves], f0s, fs, device=dev))
    off = np.concatenate(([0], np.cumsum([len(f) for f in f0s])))
    fh = feats.cpu().numpy()
    hp = SimpleNamespace(synth_fs=fs, num_coded_sps=60, num_bap=1, sp_type="mcep", do_post_filtering=False, synth_file_suffix="_t",
                         synth_ext="wav", synth_dir=str(tmp_path / "synth"), preemphasis=0.0)
    out = {"spk/a": fh[off[0]:off[1]], "b": fh[off[1]:off[2]]}
    Synthesiser.run_world_synth(out, hp)
    for name, rows in (("a", out["spk/a"]), ("b", out["b"])):
        data, rfs = AudioProcessing.read_wav(str(tmp_path / "synth" / (name + "_t_60mcep_WORLD.wav")))
        assert rfs == fs
        assert abs(len(data) / fs / 0.005 - len(rows)) < 10  # test_AcousticModelTrainer.py:162-168


def test_baseline_sized_batch_properties():
    """BASELINE config 2 shape (6.5 s, 22.05 kHz utterances), a slice of 48 of them: properties that do not need the oracle."""
    from idiaptts_b200 import ops, pipeline, synthetic
    dev = torch.device("cuda", 0)
    fs = 22050
    waves, f0s = synthetic.make_corpus(48, fs, seed=2, mean_dur=6.5, device=dev)
    x = torch.cat(waves)
    lens = np.array([w.numel() for w in waves])
    fl = np.array([len(f) for f in f0s])
    assert np.all(fl == 1301) and np.all(lens == 143325)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    batch = ops.RaggedBatch(x, up(np.concatenate(([0], np.cumsum(lens))).astype(np.int64)), up(np.concatenate(f0s)),
                            up(np.concatenate([np.arange(n) * 5.0 / 1000.0 for n in fl])),
                            up(np.concatenate(([0], np.cumsum(fl))).astype(np.int64)), up(np.repeat(np.arange(48, dtype=np.int32), fl)), fs)
    an = pipeline.WorldAnalyzer(fs, 60, device=dev, chunk_frames=20000)
    feats, sums, st = an.extract(batch)
    assert ops.raise_for_status(st, "big") == 0
    f = feats.cpu().numpy()
    assert f.shape == (48 * 1301, 64) and np.isfinite(f).all()
    # idempotence / determinism: a second pass gives identical bits
    feats2, sums2, _ = an.extract(batch)
    assert torch.equal(feats, feats2)
    # statistics are the sums of what was written
    np.testing.assert_allclose(sums.cpu().numpy()[:64], f.astype(np.float64).sum(0), rtol=1e-9, atol=1e-5)
    # vuv column equals the F0 track's voicing, lf0 is continuous and positive wherever any voiced frame exists
    assert np.array_equal(f[:, 61] > 0, np.concatenate(f0s) > 0)
    assert np.all(f[:, 60] > 3.4)
    # unvoiced frames carry the WORLD constant in every band; voiced frames have bap <= 0
    unv = f[:, 61] == 0
    assert np.all(f[unv][:, 62:] == np.float32(-8.685697e-12)) and np.all(f[:, 62:] <= 0)
    # analysis -> synthesis round trip keeps the utterance energy within a few dB
    syn = pipeline.WorldSynthesizer(fs, 60, device=dev)
    y, out_off, st = syn.synthesize(feats[:4 * 1301].contiguous(), batch.frame_off[:5].contiguous())
    assert ops.raise_for_status(st, "synth") == 0
    y = y.cpu().numpy().astype(np.float64)
    for u in range(4):
        orig = waves[u].cpu().numpy().astype(np.float64) / 32768.0
        got = y[out_off[u]:out_off[u + 1]]
        assert abs(10 * np.log10((got ** 2).mean() / (orig ** 2).mean())) < 3.0


def test_extract_from_host_matches_resident_extract():
    """The streamed host -> device -> host path (copies overlapped with the analysis of neighbouring chunks) returns exactly
    the rows and statistics of the resident-input path, also when a chunk boundary falls inside an utterance."""
    from idiaptts_b200 import ops, pipeline, synthetic
    dev = torch.device("cuda", 0)
    fs = 22050
    waves, f0s = synthetic.make_corpus(7, fs, seed=11, mean_dur=1.2, std_dur=0.3)
    batch = ops.RaggedBatch.from_host([w.numpy() for w in waves], f0s, fs, device=dev)
    an = pipeline.WorldAnalyzer(fs, 60, device=dev, chunk_frames=500)   # 7 utterances of ~240 frames: ragged chunks
    feats, sums, status = an.extract(batch)
    ops.raise_for_status(status, "resident")
    host = {"x": batch.x.cpu().pin_memory(), "so": batch.sample_off.cpu().pin_memory(), "f0": batch.f0.cpu().pin_memory(),
            "t": batch.t.cpu().pin_memory(), "fo": batch.frame_off.cpu().pin_memory(), "fu": batch.frame_utt.cpu().pin_memory()}
    out = torch.empty(feats.shape, dtype=torch.float32).pin_memory()
    bufs = None
    for _ in range(2):  # second call reuses the device buffers
        out.zero_()
        f2, s2, st2, bufs = an.extract_from_host(host, out, dev_buffers=bufs)
        torch.cuda.synchronize()
        ops.raise_for_status(st2, "streamed")
        assert torch.equal(out, feats.cpu()) and torch.equal(f2, feats)
        assert torch.allclose(s2, sums, rtol=1e-12, atol=0.0)  # fp64 atomics: summation order may differ in the last bits


def test_synthesize_corpus_without_round_trips_matches_single_calls():
    """The corpus-level synthesis (batches queued back to back: offsets derived on the device, response slab allocated whole, no
    pulse-count read-back; features from pinned host memory, waveforms to pinned host memory) returns exactly the samples of
    per-batch calls that size everything from read-back counts, also when the last batch is ragged."""
    from idiaptts_b200 import ops, pipeline, synthetic
    dev = torch.device("cuda", 0)
    fs = 22050
    waves, f0s = synthetic.make_corpus(11, fs, seed=13, mean_dur=1.0, std_dur=0.4)
    batch = ops.RaggedBatch.from_host([w.numpy() for w in waves], f0s, fs, device=dev)
    an = pipeline.WorldAnalyzer(fs, 60, device=dev)
    feats, _, st = an.extract(batch)
    assert ops.raise_for_status(st, "extract") & ~8 == 0
    fo = batch.frame_off.cpu().numpy()
    syn = pipeline.WorldSynthesizer(fs, 60, device=dev)
    # reference: one call per batch of 4 utterances through the read-back path
    ref = []
    for u0 in range(0, 11, 4):
        u1 = min(11, u0 + 4)
        off = torch.from_numpy(fo[u0:u1 + 1] - fo[u0]).to(dev)
        y, _, s = syn.synthesize(feats[fo[u0]:fo[u1]], off)
        assert ops.raise_for_status(s, "synth") == 0
        ref.append(y)
    ref = torch.cat(ref)
    feats_host = feats.cpu().pin_memory()
    y_host = torch.zeros(ref.numel(), dtype=torch.float32).pin_memory()
    y_dev = torch.zeros(ref.numel(), dtype=torch.float32, device=dev)
    out_off, s = syn.synthesize_corpus(None, fo, batch_utts=4, feats_host=feats_host, out_host=y_host, out=y_dev)
    torch.cuda.synchronize()
    assert ops.raise_for_status(s, "synthesize_corpus") == 0
    assert int(out_off[-1]) == ref.numel()
    assert torch.equal(y_dev, ref) and torch.equal(y_host, ref.cpu())


@pytest.mark.parametrize("add_deltas", [False, True])
def test_gen_data_pieces_equal_one_piece(tmp_path, add_deltas):
    """gen_data's read / extract / write pipeline: splitting the shard into pieces (here one or two utterances each, so the two
    buffers of every hand-over are reused several times) leaves the same bytes on disk and the same statistics as one piece;
    errors raised by the reader and writer threads reach the caller."""
    import filecmp
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    fs = 22050
    ids, cache, waves, f0s = _write_corpus(tmp_path, 7, fs, 0.6, seed=11)
    one = WorldFeatLabelGen(str(tmp_path / "one"), add_deltas=add_deltas, num_coded_sps=60, num_bap=2, f0_cache=cache)
    ld1, m1, s1 = one.gen_data(str(tmp_path / "wav"), str(tmp_path / "one"), file_id_list="train.txt", id_list=ids, return_dict=True)
    many = WorldFeatLabelGen(str(tmp_path / "many"), add_deltas=add_deltas, num_coded_sps=60, num_bap=2, f0_cache=cache,
                             io_chunk_seconds=1.0, io_threads=2)
    for rep in range(2):                                       # the second call reuses the pinned buffers of the first
        ld2, m2, s2 = many.gen_data(str(tmp_path / "wav"), str(tmp_path / "many"), file_id_list="train.txt", id_list=ids, return_dict=True)
        assert list(ld2) == ids
        for i in ids:
            assert np.array_equal(ld1[i], ld2[i])
            for d in ("mcep60", "lf0", "vuv", "bap"):
                assert filecmp.cmp(str(tmp_path / "one" / d / (i + ".npz")), str(tmp_path / "many" / d / (i + ".npz")), shallow=False)
        if add_deltas:
            for a, b in zip(m1, m2):
                np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)
            for a, b in zip(s1, s2):
                np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10)
        else:
            np.testing.assert_allclose(m1, m2, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(s1, s2, rtol=1e-8, atol=1e-10)
    # statistics only (no directory, no dictionary): nothing is copied back or written
    m3, s3 = many.gen_data(str(tmp_path / "wav"), None, id_list=ids)
    if not add_deltas:
        np.testing.assert_allclose(m1, m3, rtol=1e-10, atol=1e-12)
    # reader thread: a cached F0 of the wrong length; writer thread: an output directory that went away
    bad = dict(cache)
    bad[ids[4]] = cache[ids[4]][:-1]
    with pytest.raises(ValueError, match="cached F0"):
        many.gen_data(str(tmp_path / "wav"), str(tmp_path / "many"), id_list=ids, f0_cache=bad)
    import shutil

    class Vanishing(WorldFeatLabelGen):
        def _create_directories(self, dir_out):
            super()._create_directories(dir_out)
            shutil.rmtree(os.path.join(dir_out, "bap"))
    v = Vanishing(str(tmp_path / "gone"), add_deltas=add_deltas, num_coded_sps=60, num_bap=2, f0_cache=cache, io_chunk_seconds=1.0)
    with pytest.raises(ValueError, match="cannot write"):
        v.gen_data(str(tmp_path / "wav"), str(tmp_path / "gone"), id_list=ids)
    # and the generator is usable afterwards
    ld4, _, _ = many.gen_data(str(tmp_path / "wav"), None, id_list=ids[:2], return_dict=True)
    assert np.array_equal(ld4[ids[1]], ld1[ids[1]])


def test_gen_data_other_sample_widths(tmp_path):
    """32-bit PCM files take the general per-file reader (float64 samples) through the same pipeline."""
    import wave
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    fs = 16000
    ids, cache, waves, f0s = _write_corpus(tmp_path, 3, fs, 0.5, seed=12)
    os.makedirs(str(tmp_path / "wav32"))
    for i, w in zip(ids, waves):
        with wave.open(str(tmp_path / "wav32" / (i + ".wav")), "wb") as f:
            f.setnchannels(1)
            f.setsampwidth(4)
            f.setframerate(fs)
            f.writeframes((w.numpy().astype(np.int32) << 16).tobytes())
    a, _, _ = WorldFeatLabelGen(None, num_coded_sps=60, f0_cache=cache, mgc_alpha=0.58).gen_data(str(tmp_path / "wav"), None, id_list=ids,
                                                                                                 return_dict=True)
    b, _, _ = WorldFeatLabelGen(None, num_coded_sps=60, f0_cache=cache, mgc_alpha=0.58, io_chunk_seconds=0.6).gen_data(
        str(tmp_path / "wav32"), None, id_list=ids, return_dict=True)
    for i in ids:
        assert a[i].shape == b[i].shape and np.array_equal(a[i][:, 61], b[i][:, 61])
        assert np.abs(a[i][:, :60] - b[i][:, :60]).max() < 1e-3
