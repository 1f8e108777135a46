"""WORLD synthesis on the GPU against the oracle (PARITY UNPINNED by reference goldens: the oracle's synthesis half is a
restatement).  Tolerance: resynthesised waveform SNR > 60 dB given identical features."""
import numpy as np
import pytest
import torch

from conftest import golden_utterance
from oracle import glue_np, world_np

pytestmark = pytest.mark.gpu


def snr_db(ref, out):
    return 10 * np.log10((ref ** 2).sum() / max(((out - ref) ** 2).sum(), 1e-300))


def test_randn_table_is_worlds_stream():
    from idiaptts_b200 import ops
    tab = ops.randn_table(70000, torch.device("cuda", 0)).cpu().numpy()
    ref = world_np.xorshift_randn_sequence(70000)
    assert np.array_equal(tab[:70000], ref)  # jump-ahead reproduces the sequential generator bit for bit
    np.testing.assert_allclose(tab[:4], [-1.32764, -0.622855, -1.609181, 1.179765], atol=1e-6)


@pytest.mark.parametrize("id_", ["LJ001-0008", "LJ001-0002"])
def test_synthesize_vs_oracle_16k(golden, id_):
    from idiaptts_b200.compat import pyworld as pw
    x, c, f0, fs = golden_utterance(golden, id_)
    t = world_np.temporal_positions(len(f0))
    sp = world_np.cheaptrick(x, f0, t, fs)
    ap = world_np.d4c(x, f0, t, fs)
    y_ref = world_np.synthesize(f0, sp, ap, fs)
    y = pw.synthesize(f0, sp, ap, fs)
    assert len(y) == len(y_ref) == int(len(f0) * 5.0 * fs / 1000)
    assert snr_db(y_ref, y) > 100  # tolerance 60 dB


def test_pulse_positions_bit_exact_with_leading_silence():
    """At 16 kHz the 500 Hz unvoiced default puts the phase EXACTLY on multiples of 2 pi every 32 samples: pulse positions are
    decided by the rounding of the sequential phase accumulation, which the GPU path therefore reproduces exactly."""
    from idiaptts_b200 import ops
    dev = torch.device("cuda", 0)
    fs, T = 16000, 400
    f0 = np.zeros(T)
    f0[150:300] = np.linspace(110.0, 180.0, 150)
    rng = np.random.default_rng(0)
    sp = np.abs(rng.standard_normal((T, 513))) * 1e-3 + 1e-4
    ap = np.clip(rng.uniform(0.05, 0.9, (T, 513)), 0.001, 0.999)
    y_ref = world_np.synthesize(f0, sp, ap, fs)
    foff = torch.tensor([0, T], dtype=torch.int64, device=dev)
    y, _, st = ops.synthesize(torch.from_numpy(f0).to(dev), torch.from_numpy(sp).to(dev), torch.from_numpy(ap).to(dev), foff, fs)
    assert ops.raise_for_status(st, "synth") == 0
    assert snr_db(y_ref, y.cpu().numpy()) > 100


def test_feature_domain_round_trip_22k_batch():
    """Synthesiser.run_world_synth path at 22.05 kHz (nap = 2) for a ragged batch: every utterance equals the oracle's
    world_features_to_raw and equals its own single-utterance synthesis."""
    from idiaptts_b200 import ops, pipeline, synthetic
    dev = torch.device("cuda", 0)
    fs = 22050
    waves, f0s = synthetic.make_corpus(3, fs, seed=4, mean_dur=0.8, std_dur=0.3)
    an = pipeline.WorldAnalyzer(fs, 60, device=dev)
    batch = ops.RaggedBatch.from_host([w.numpy() for w in waves], f0s, fs, device=dev)
    feats, _, st = an.extract(batch)
    ops.raise_for_status(st, "extract")
    syn = pipeline.WorldSynthesizer(fs, 60, device=dev)
    y, out_off, st = syn.synthesize(feats, batch.frame_off)
    ops.raise_for_status(st, "synth")
    y = y.cpu().numpy().astype(np.float64)
    fh = feats.cpu().numpy()
    foff = batch.frame_off.cpu().numpy()
    for u in range(3):
        rows = fh[foff[u]:foff[u + 1]]
        amp = glue_np.mcep_to_amp_sp(rows[:, :60], fs, alpha=an.alpha)
        ref = glue_np.world_features_to_raw(amp, rows[:, 60].copy(), rows[:, 61].copy(), rows[:, 62:].copy(), fs)
        got = y[out_off[u]:out_off[u + 1]]
        assert len(got) == len(ref) == int(len(rows) * 5.0 * fs / 1000)
        assert snr_db(ref, got) > 60
        single, _, _ = syn.synthesize(torch.from_numpy(rows).to(dev), torch.tensor([0, len(rows)], dtype=torch.int64, device=dev))
        assert np.array_equal(single.cpu().numpy().astype(np.float64), got)  # batching is bit-reproducible (no atomics)


def test_world_features_to_raw_and_deemphasis(golden):
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    T = len(f0)
    sp = world_np.cheaptrick(x, f0, world_np.temporal_positions(T), fs)
    amp = np.sqrt(sp)
    lf0, vuv = c[:, 60].copy(), c[:, 63].copy()
    bap = c[:, 64:65].copy()
    ref = glue_np.world_features_to_raw(amp, lf0.copy(), vuv.copy(), bap, fs, preemphasis_coef=0.97)
    out = WorldFeatLabelGen.world_features_to_raw(amp, lf0.copy(), vuv.copy(), bap, fs, preemphasis=0.97)
    assert out.dtype == np.float64 and len(out) == len(ref)
    assert snr_db(ref, out) > 60


def test_fast_render_path_vs_f64_path_and_oracle(golden):
    """The batched fast path (one warp per pulse, single-precision transforms, float32 responses) against the fp64 path and the
    oracle: same sample count, identical pulse table, SNR > 100 dB against the fp64 kernel (tolerance 60 dB against the oracle),
    incl. the edge cases: leading / trailing silence (exact 500 Hz ties), a voiced-but-aperiodic frame, one-frame utterances."""
    from idiaptts_b200 import ops
    dev = torch.device("cuda", 0)
    x, c, f0, fs = golden_utterance(golden, "LJ001-0002")
    T = len(f0)
    t = world_np.temporal_positions(T)
    sp = world_np.cheaptrick(x, f0, t, fs)
    ap = world_np.d4c(x, f0, t, fs)
    ap[200:210] = 1.0 - 1e-12                 # voiced f0 with an all-aperiodic spectrum: ar[0] > 0.999 -> no periodic response
    y_ref = world_np.synthesize(f0, sp, ap, fs)
    # ragged batch: the utterance, a 1-frame utterance, an all-unvoiced one
    f0b = np.concatenate((f0, f0[300:301], np.zeros(50)))
    spb = np.concatenate((sp, sp[300:301], sp[:50]))
    apb = np.concatenate((ap, ap[300:301], ap[:50]))
    foff = torch.tensor([0, T, T + 1, T + 51], dtype=torch.int64, device=dev)
    args = (torch.from_numpy(f0b).to(dev), torch.from_numpy(spb).to(dev), torch.from_numpy(apb).to(dev))
    out = {}
    for precision in ("f64", "fast"):
        dbg = {}
        plan = ops.synth_timebase(args[0], foff, fs, 1024)
        y, out_off, st = ops.synth_render(plan, args[1], args[2], precision=precision, debug=dbg)
        assert ops.raise_for_status(st, "synth") == 0
        out[precision] = (y.cpu().numpy(), out_off, dbg)
    y64, off64, d64 = out["f64"]
    y32, off32, d32 = out["fast"]
    assert np.array_equal(off64, off32) and np.array_equal(d64["num_pulses"], d32["num_pulses"])
    assert d32["response"].dtype == torch.float32 and d64["response"].dtype == torch.float64
    for u in range(3):
        a, b = y64[off64[u]:off64[u + 1]], y32[off32[u]:off32[u + 1]]
        if (a ** 2).sum() > 0:
            assert snr_db(a, b) > 100, (u, snr_db(a, b))
        else:
            assert not b.any()
    assert snr_db(y_ref, y32[:off32[1]]) > 60
    # per-pulse responses: every row the fp64 kernel wrote is reproduced (relative to the row's own energy)
    r64, r32 = d64["response"].cpu().numpy(), d32["response"].cpu().numpy().astype(np.float64)
    npul, poff = d64["num_pulses"], d64["pulse_off"]
    worst = 1e9
    for u in range(3):
        for p in range(int(npul[u])):
            a, b = r64[poff[u] + p], r32[poff[u] + p]
            if (a ** 2).sum() == 0:   # the last pulse of an utterance has no noise segment (noise_size = 0): an all-zero response
                assert not b.any()
                continue
            worst = min(worst, snr_db(a, b))
    assert worst > 90, worst


@pytest.mark.parametrize("precision", ["fast", "f64"])
def test_reference_resynthesis_criterion_on_all_fixtures(golden, precision):
    """The only constraint the reference itself holds on the synthesis half (test/integration/data_preparation/world/
    test_WorldFeatLabelGen.py:761-763): wav -> world features -> raw, both peak-normalised, sum of squared errors < 10000 --
    run on the GPU path (analysis AND synthesis, pre-emphasis 0.97) over all 9 fixture utterances."""
    from idiaptts_b200 import ops, pipeline
    dev = torch.device("cuda", 0)
    fs = 16000
    ids = ["LJ001-%04d" % i for i in range(1, 10)]
    waves = [golden[i + "/wav"] for i in ids]
    f0s = []
    for i in ids:
        c = golden[i + "/cmp"]
        f0s.append(np.where(c[:, 63] > 0, np.exp(c[:, 60].astype(np.float64)), 0.0))
    batch = ops.RaggedBatch.from_host(waves, f0s, fs, preemphasis=0.97, device=dev)
    an = pipeline.WorldAnalyzer(fs, 60, mgc_alpha=0.58, device=dev)
    feats, _, st = an.extract(batch)
    assert ops.raise_for_status(st, "extract") & ~8 == 0
    syn = pipeline.WorldSynthesizer(fs, 60, mgc_alpha=0.58, device=dev, precision=precision)
    y, out_off, st = syn.synthesize(feats, batch.frame_off, preemphasis=0.97)
    assert ops.raise_for_status(st, "synth") == 0
    y = y.cpu().numpy().astype(np.float64)
    for u, w in enumerate(waves):
        raw = w.astype(np.float64) / 32768.0
        rec = y[out_off[u]:out_off[u + 1]]
        n = min(len(raw), len(rec))
        assert abs(len(raw) - len(rec)) < 10 * 80          # length within 10 frames (test_AcousticModelTrainer.py:162-168)
        a, b = raw[:n], rec[:n] / np.abs(rec[:n]).max()          # the reference scales only the reconstruction to [-1, 1]
        assert ((a - b) ** 2).sum() < 10000, ids[u]
