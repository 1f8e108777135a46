"""GPU parity of the F0 stage: b2w_dio + b2w_stonemask against the reference's golden vectors (the lf0 / vuv
columns of fixtures/WORLD/cmp_mcep20/*.cmp, produced by pyworld.wav2world) and against the numpy oracle (oracle/dio_np.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_utterance
from oracle import dio_np, glue_np, world_np

pytestmark = pytest.mark.gpu

IDS = ["LJ001-%04d" % i for i in range(1, 10)]


def _batch(waves, fs, preemphasis=0.0, frame_period=5.0):
    from idiaptts_b200 import ops
    f0s = [np.zeros(world_np.num_frames(len(w), fs, frame_period)) for w in waves]
    return ops.RaggedBatch.from_host(waves, f0s, fs, frame_period=frame_period, preemphasis=preemphasis, device="cuda")


def test_f0_reproduces_reference_fixtures(golden):
    """All 9 reference utterances as one ragged int16 batch, pre-emphasis applied on load: vuv bit-exact, lf0 to float32 round-off."""
    from idiaptts_b200 import ops
    waves = [golden[i + "/wav"] for i in IDS]
    batch = _batch(waves, 16000, preemphasis=0.97)
    f0 = ops.estimate_f0(batch).cpu().numpy()
    off = batch.frame_off.cpu().numpy()
    total = 0
    for u, id_ in enumerate(IDS):
        c = golden[id_ + "/cmp"]
        f = f0[off[u]:off[u + 1]]
        assert len(f) == c.shape[0]
        assert np.array_equal(f > 0, c[:, 63] > 0), id_
        v = f > 0
        assert np.abs(np.log(f[v]).astype(np.float32) - c[v, 60]).max() < 1e-6, id_
        total += len(f)
    assert total == 11579
    # ... and the label columns the pipeline derives from it (interpolate_lin, WorldFeatLabelGen.py:798-802)
    lf0, vuv = ops.lf0_vuv(batch.f0, batch.frame_off)
    lf0, vuv = lf0.cpu().numpy(), vuv.cpu().numpy()
    for u, id_ in enumerate(IDS):
        c = golden[id_ + "/cmp"]
        assert np.array_equal(vuv[off[u]:off[u + 1], 0], c[:, 63])
        assert np.abs(lf0[off[u]:off[u + 1], 0] - c[:, 60]).max() < 2e-6


@pytest.mark.parametrize("fs,step2", [(22050, "erosion"), (22050, "sections"), (48000, "erosion"), (16000, "erosion")])
def test_dio_and_stonemask_match_oracle(fs, step2):
    from idiaptts_b200 import ops, synthetic
    waves, _ = synthetic.make_corpus(3, fs, seed=31 + fs % 7, mean_dur=1.2, device="cpu")
    waves = [w.numpy() for w in waves]
    batch = _batch(waves, fs)
    f0_dio = ops.dio(batch, step2=step2).cpu().numpy()
    off = batch.frame_off.cpu().numpy()
    refs = []
    for u, w in enumerate(waves):
        x = w.astype(np.float64) / 32768.0
        T = off[u + 1] - off[u]
        t = np.arange(T) * 5.0 / 1000.0  # WORLD: i * frame_period / 1000 (not i * 0.005: StoneMask rounds (t + dt) * fs to samples)
        cands, scores = dio_np.dio_candidates(x, fs, t)
        ref = dio_np.fix_f0_contour(5.0, cands, dio_np.best_f0_contour(cands, scores), 71.0, 0.1, step2=step2)
        got = f0_dio[off[u]:off[u + 1]]
        assert np.array_equal(got > 0, ref > 0)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=0)
        refs.append(dio_np.stonemask(x, ref, t, fs))
    assert (f0_dio > 0).mean() > 0.2  # the synthetic corpus is ~65 % voiced
    f0 = ops.stonemask(batch, torch.from_numpy(f0_dio).cuda()).cpu().numpy()
    ref = np.concatenate(refs)
    assert np.array_equal(f0 > 0, ref > 0)
    np.testing.assert_allclose(f0, ref, rtol=1e-9, atol=0)


def test_dio_chunking_short_and_empty_utterances():
    from idiaptts_b200 import ops, synthetic
    fs = 16000
    waves, _ = synthetic.make_corpus(4, fs, seed=5, mean_dur=0.9, device="cpu")
    waves = [w.numpy() for w in waves]
    waves.insert(1, np.zeros(400, np.int16))          # 6 frames <= voice_range_minimum: FixF0Contour leaves it unvoiced
    waves.insert(3, np.zeros(3000, np.int16))         # digital silence: no zero crossings at all
    batch = _batch(waves, fs)
    a = ops.estimate_f0(batch).cpu().numpy()
    b = ops.stonemask(batch, ops.dio(batch, max_chunk_samples=20000)).cpu().numpy()  # every utterance its own chunk
    assert np.array_equal(a, b)
    off = batch.frame_off.cpu().numpy()
    assert not a[off[1]:off[2]].any() and not a[off[3]:off[4]].any()
    for u in (0, 2, 5):
        x = waves[u].astype(np.float64) / 32768.0
        ref, _ = dio_np.wav2world_f0(x, fs)
        assert np.array_equal(a[off[u]:off[u + 1]] > 0, ref > 0)
        np.testing.assert_allclose(a[off[u]:off[u + 1]], ref, rtol=1e-9)


def test_pyworld_compatible_f0_calls(golden):
    from idiaptts_b200.compat import pyworld
    x, c, _, fs = golden_utterance(golden, "LJ001-0002")
    _f0, t = pyworld.dio(x, fs)
    assert _f0.dtype == np.float64 and t.dtype == np.float64 and len(_f0) == len(t) == c.shape[0]
    assert np.array_equal(t, np.arange(len(t)) * 5.0 / 1000.0)
    f0 = pyworld.stonemask(x, _f0, t, fs)
    assert np.array_equal(f0 > 0, c[:, 63] > 0)
    f0b, sp, ap = pyworld.wav2world(x, fs)
    assert np.array_equal(f0, f0b) and sp.shape == (len(f0), 513) and ap.shape == sp.shape
    with pytest.raises(ValueError):
        pyworld.dio(x, fs, speed=4)


def test_world_extract_features_without_cached_f0(golden):
    """The reference's call (WorldFeatLabelGen.world_extract_features(raw, fs, hop), :779-807) with no F0 supplied."""
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    x, c, f0g, fs = golden_utterance(golden, "LJ001-0008")
    amp_sp, lf0, vuv, bap = WorldFeatLabelGen.world_extract_features(x, fs, 5)
    assert np.array_equal(vuv[:, 0], c[:, 63])
    assert np.abs(lf0[:, 0] - c[:, 60]).max() < 2e-6
    assert np.abs(bap[:, 0] - c[:, 64]).max() < 3e-5
    from idiaptts_b200.AudioProcessing import AudioProcessing
    mc = AudioProcessing.extract_mcep(amp_sp, num_coded_sps=20, mgc_alpha=0.58)
    assert glue_np.mcd_db(c[:, :20], mc) < 1e-3


@pytest.mark.parametrize("add_deltas", [False, True])
def test_lf0labelgen_gen_data(tmp_path, golden, add_deltas):
    """LF0LabelGen.gen_data (LF0LabelGen.py:212-321) on three reference utterances against the oracle path of the same code."""
    import wave
    from idiaptts_b200.LF0LabelGen import LF0LabelGen
    ids = IDS[1:2] + IDS[7:9]
    os.makedirs(str(tmp_path / "wav"))
    for i in ids:
        with wave.open(str(tmp_path / "wav" / (i + ".wav")), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
            w.writeframes(golden[i + "/wav"].astype(np.int16).tobytes())
    gen = LF0LabelGen(str(tmp_path / "out"), add_deltas=add_deltas)
    label_dict, mean, std = gen.gen_data(str(tmp_path / "wav"), str(tmp_path / "out"), file_id_list="train.txt", id_list=ids,
                                         add_deltas=add_deltas, return_dict=True)
    allrows = []
    for i in ids:
        x = golden[i + "/wav"].astype(np.float64) / 32768.0   # no pre-emphasis here (soundfile.read, :262)
        f0, _ = dio_np.wav2world_f0(x, 16000)
        lf0, vuv = glue_np.interpolate_lin(glue_np.lf0_from_f0(f0, f0_silence_threshold=20))
        got = label_dict[i]
        assert got.shape == (len(f0), 4 if add_deltas else 2)
        assert np.array_equal(got[:, -1], vuv[:, 0])
        assert np.abs(got[:, 0] - lf0[:, 0]).max() < 2e-6
        if add_deltas:
            assert np.array_equal(got[:, 1:2], glue_np.compute_deltas(got[:, 0:1]))
            assert np.array_equal(got[:, 2:3], glue_np.compute_deltas(got[:, 1:2]))
        if add_deltas:  # the reference writes all four columns into <id>.lf0_deltas (:285)
            raw = np.fromfile(str(tmp_path / "out" / "lf0" / (i + ".lf0_deltas")), dtype=np.float32).reshape(-1, 4)
            assert np.array_equal(raw, got)
        else:
            assert np.array_equal(LF0LabelGen.load_sample(i, str(tmp_path / "out")), got)
        allrows.append(got)
    allrows = np.concatenate(allrows).astype(np.float64)
    np.testing.assert_allclose(mean[0], allrows[:, 0].mean(), rtol=1e-5)
    assert mean[-1] == 0.0 and abs(std[-1] - 1.0) < 1e-12
    params = gen.get_normalisation_params(str(tmp_path / "out"), "train")
    assert params[0].shape[-1] == (4 if add_deltas else 2)
    norm = gen.preprocess_sample(label_dict[ids[0]])
    np.testing.assert_allclose(gen.postprocess_sample(norm), label_dict[ids[0]], atol=1e-5)


def test_f0_scale_invariance_at_corpus_scale():
    """Size-independent property at the benchmark's utterance shape (6.5 s @ 22.05 kHz): every DIO / StoneMask decision is
    homogeneous in the waveform, and scaling by a power of two is exact in IEEE arithmetic, so DIO(x / 4) == DIO(x) bit for bit."""
    from idiaptts_b200 import ops, synthetic
    fs = 22050
    waves, _ = synthetic.make_corpus(48, fs, seed=77, mean_dur=6.5, device="cpu")
    xs = [w.numpy().astype(np.float64) / 32768.0 for w in waves]
    ba, bb = _batch(xs, fs), _batch([x * 0.25 for x in xs], fs)
    da, db = ops.dio(ba), ops.dio(bb)
    assert torch.equal(da, db)                      # DIO: exactly homogeneous
    a, b = ops.stonemask(ba, da).cpu().numpy(), ops.stonemask(bb, db).cpu().numpy()
    assert np.array_equal(a > 0, b > 0)             # StoneMask adds 1e-12 to an amplitude sum (FixF0): homogeneous to ~1e-12
    np.testing.assert_allclose(a, b, rtol=1e-9)
    assert 0.3 < (a > 0).mean() < 0.9
    # ... and an utterance's track does not depend on what else is in the batch
    c = ops.estimate_f0(_batch(xs[5:6], fs)).cpu().numpy()
    off = np.concatenate(([0], np.cumsum([world_np.num_frames(len(x), fs) for x in xs])))
    assert np.array_equal(c, a[off[5]:off[6]])
