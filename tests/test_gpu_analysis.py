"""Parity of the CUDA analysis kernels (through the C ABI) against the reference's golden vectors and the CPU oracle.

Tolerances (BASELINE.json north_star): frame counts and vuv bit-exact; spectral envelope relative error <= 1e-4;
mel-cepstral distortion < 0.01 dB; the tighter numbers asserted here are what the kernels actually deliver."""
import numpy as np
import pytest
import torch

from conftest import golden_utterance
from oracle import glue_np, sptk_np, world_np

pytestmark = pytest.mark.gpu

IDS = ["LJ001-%04d" % i for i in range(1, 10)]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


@pytest.mark.parametrize("id_", ["LJ001-0002", "LJ001-0008"])
def test_cheaptrick_vs_oracle(golden, id_):
    from idiaptts_b200.compat import pyworld as pw
    x, c, f0, fs = golden_utterance(golden, id_)
    t = world_np.temporal_positions(len(f0))
    ref = world_np.cheaptrick(x, f0, t, fs)
    out = pw.cheaptrick(x, f0, t, fs)
    assert out.shape == ref.shape and out.dtype == np.float64
    assert (np.abs(out - ref) / ref).max() < 1e-6  # tolerance 1e-4


@pytest.mark.parametrize("id_", IDS)
def test_fused_analysis_reproduces_reference_cmp(golden, dev, id_):
    """int16 wav + in-kernel pre-emphasis 0.97 -> CheapTrick (float32 plane) -> mcep20, D4C -> bap, against the reference's
    own cmp_mcep20 fixtures (the same comparison that pins the oracle)."""
    from idiaptts_b200 import ops
    c = golden[id_ + "/cmp"]
    f0 = np.where(c[:, 63] > 0, np.exp(c[:, 60].astype(np.float64)), 0.0)
    wav = golden[id_ + "/wav"]
    assert ops.num_frames(len(wav), 16000) == c.shape[0]  # frame count: bit-exact
    batch = ops.RaggedBatch.from_host([wav], [f0], 16000, preemphasis=0.97, device=dev)
    sp, status = ops.cheaptrick(batch, out_dtype=torch.float32)
    mc, _ = ops.mcep(sp, 19, 0.58, is_power=True, status=status)
    coarse, voiced, _ = ops.d4c_coarse(batch, status=status)
    bap = ops.bap_from_coarse(coarse, voiced, 16000, 1024)
    assert ops.raise_for_status(status, "analysis") & ~8 == 0
    mc = mc.cpu().numpy()
    assert np.abs(mc - c[:, :20]).max() < 2e-5
    assert glue_np.mcd_db(c[:, :20], mc) < 1e-4  # tolerance 0.01 dB
    bap = bap.cpu().numpy()
    assert np.abs(bap[:, 0] - c[:, 64]).max() < 5e-4  # dB; the fast D4C path runs its FFTs in single precision (3e-5 with precision="f64")
    assert np.all(bap[c[:, 63] == 0, 0] == np.float32(-8.685697e-12))


def test_d4c_and_codec_vs_oracle(golden):
    from idiaptts_b200.compat import pyworld as pw
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    t = world_np.temporal_positions(len(f0))
    ap_ref = world_np.d4c(x, f0, t, fs)
    ap = pw.d4c(x, f0, t, fs)
    unv_ref = ap_ref[:, 0] > 0.99
    assert np.array_equal(ap[:, 0] > 0.99, unv_ref)  # LoveTrain decisions identical
    assert np.abs(ap - ap_ref).max() < 1e-9
    bap = pw.code_aperiodicity(ap, fs)
    np.testing.assert_allclose(bap, world_np.code_aperiodicity(ap_ref, fs), atol=1e-8)
    np.testing.assert_allclose(pw.decode_aperiodicity(bap, fs, 1024), world_np.decode_aperiodicity(bap, fs, 1024), atol=1e-12)
    with pytest.raises(ValueError):
        pw.decode_aperiodicity(np.zeros((3, 2)), 16000, 1024)  # wrong band count for fs


def test_cheaptrick_fast_path_vs_oracle(golden, dev):
    """The float32 plane of the fused path (warp-per-frame mixed-precision kernel: b2w_cheaptrick with B2W_F32 at fft size 1024)
    against the fp64 oracle on ALL 11 579 frames of the 9 reference utterances (pre-emphasis 0.97, int16 input): spectral
    envelope relative error <= 1e-4 (north_star tolerance), and against the fp64 kernel; edge cases: frames whose window hangs
    over both ends of a short utterance, f0 below the floor (500 Hz default window), the highest f0 the track can carry."""
    from idiaptts_b200 import ops
    worst = 0.0
    for id_ in IDS:
        c = golden[id_ + "/cmp"]
        f0 = np.where(c[:, 63] > 0, np.exp(c[:, 60].astype(np.float64)), 0.0)
        wav = golden[id_ + "/wav"]
        batch = ops.RaggedBatch.from_host([wav], [f0], 16000, preemphasis=0.97, device=dev)
        sp32, st = ops.cheaptrick(batch, out_dtype=torch.float32)
        sp64, st = ops.cheaptrick(batch, out_dtype=torch.float64, status=st)
        assert ops.raise_for_status(st, "cheaptrick") == 0
        rel = ((sp32.double() - sp64) / sp64).abs().max().item()
        worst = max(worst, rel)
    assert worst < 1e-4, worst
    x, c, f0, fs = golden_utterance(golden, "LJ001-0003")
    t = world_np.temporal_positions(len(f0))
    ref = world_np.cheaptrick(x, f0, t, fs)
    batch = ops.RaggedBatch.from_host([x], [f0], fs, device=dev)
    sp32, _ = ops.cheaptrick(batch, out_dtype=torch.float32)
    assert (np.abs(sp32.cpu().numpy().astype(np.float64) - ref) / ref).max() < 1e-4
    # edge cases on a 0.1 s utterance
    xs = x[8000:9600]
    f0e = np.array([0.0, 40.0, 64.0, 71.0, 120.0, 250.0, 400.0, 799.0, 1000.0] + [150.0] * 12)
    te = world_np.temporal_positions(len(f0e))
    refe = world_np.cheaptrick(xs, f0e, te, fs)
    be = ops.RaggedBatch.from_host([xs], [f0e], fs, device=dev)
    spe, st = ops.cheaptrick(be, out_dtype=torch.float32)
    assert ops.raise_for_status(st, "cheaptrick") == 0
    assert (np.abs(spe.cpu().numpy().astype(np.float64) - refe) / refe).max() < 1e-4


def test_d4c_fast_path_vs_f64_path(golden, dev):
    """b2w_d4c_coarse (single-precision FFTs) against b2w_d4c_coarse_f64 on all 9 reference utterances: LoveTrain decisions
    identical, coarse aperiodicity within 5e-4 dB (a numpy emulation with single-precision pocketfft transforms shows the same
    1-2e-4 dB worst case on these utterances: it is the precision of the transform, not of this kernel); and the guard band: with the threshold moved onto a frame's own LoveTrain
    ratio the fast pass must hand that frame to the fp64 kernel, which then decides exactly like the fp64 path."""
    from idiaptts_b200 import ops
    worst = 0.0
    for id_ in IDS:
        x, c, f0, fs = golden_utterance(golden, id_)
        batch = ops.RaggedBatch.from_host([x], [f0], fs, device=dev)
        c64, v64, st = ops.d4c_coarse(batch, precision="f64")
        c32, v32, st = ops.d4c_coarse(batch, status=st, precision="fast")
        assert torch.equal(v64, v32), id_
        assert int(v32.max().item()) <= 1                         # no frame is left marked "undecided"
        m = v64.bool()
        worst = max(worst, (c64[m] - c32[m]).abs().max().item())
        assert ops.raise_for_status(st, "d4c") == 0
    assert worst < 5e-4, worst
    # guard band: thresholds equal to (and a hair around) the fp64 LoveTrain ratio of one frame
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    t = world_np.temporal_positions(len(f0))
    ap0 = world_np.d4c_lovetrain(x, fs, f0, t)
    i = int(np.argmin(np.where(f0 > 0, np.abs(ap0 - 0.85), 1.0)))
    batch = ops.RaggedBatch.from_host([x], [f0], fs, device=dev)
    for thr in (ap0[i], np.nextafter(ap0[i], 0.0), np.nextafter(ap0[i], 1.0), ap0[i] - 3e-7, ap0[i] + 3e-7):
        c64, v64, _ = ops.d4c_coarse(batch, threshold=float(thr), precision="f64")
        c32, v32, _ = ops.d4c_coarse(batch, threshold=float(thr), precision="fast")
        assert torch.equal(v64, v32), thr
        if bool(v64[i]):  # the re-evaluated frame carries the fp64 kernel's values
            assert torch.equal(c64[i], c32[i])


@pytest.mark.parametrize("order,alpha", [(19, 0.58), (59, 0.58), (59, 0.41), (79, 0.41), (39, 0.42), (24, 0.35)])
def test_mcep_vs_oracle(golden, order, alpha):
    from idiaptts_b200.compat import pysptk as ps
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    sp = world_np.cheaptrick(x, f0, world_np.temporal_positions(len(f0)), fs)
    amp = np.sqrt(sp)
    ref = [sptk_np.mcep_frame(a, order, alpha, eps=1e-8) for a in amp]
    rm = np.stack([r[0] for r in ref])
    out = ps.mcep(amp, order=order, alpha=alpha, eps=1.0e-8, min_det=0.0, etype=1, itype=3)
    assert out.shape == rm.shape
    assert np.abs(out - rm).max() < 2e-5
    assert glue_np.mcd_db(rm, out) < 1e-4
    one = ps.mcep(amp[40], order=order, alpha=alpha, eps=1.0e-8, etype=1, itype=3)  # 1-D input = one frame
    np.testing.assert_allclose(one, out[40], atol=1e-6)
    # power-spectrum input (itype=4) gives the same result as amplitude input
    out4 = ps.mcep(sp, order=order, alpha=alpha, eps=1.0e-8, etype=1, itype=4)
    assert np.abs(out4 - out).max() < 2e-5


def test_mcep_iteration_counts_and_errors(golden, dev):
    from idiaptts_b200 import ops
    from idiaptts_b200.compat import pysptk as ps
    x, c, f0, fs = golden_utterance(golden, "LJ001-0002")
    sp = world_np.cheaptrick(x, f0, world_np.temporal_positions(len(f0)), fs)
    ref_it = np.array([sptk_np.mcep_frame(np.sqrt(a), 59, 0.58, eps=1e-8)[1] for a in sp])
    iters = torch.zeros(len(sp), dtype=torch.int32, device=dev)
    ops.mcep(torch.from_numpy(sp).to(dev), 59, 0.58, is_power=True, iters=iters)
    assert np.array_equal(iters.cpu().numpy(), ref_it)  # same Newton trajectory as SPTK
    with pytest.raises(RuntimeError, match="periodogram"):
        ps.mcep(np.zeros((2, 513)), order=19, alpha=0.4, etype=0, itype=3)  # pysptk raises on zeros without eps
    with pytest.raises(ValueError):
        ps.mcep(np.ones((2, 513)), order=19, alpha=0.4, etype=1, eps=-1.0, itype=3)


def test_mc2sp_and_reference_reconstruction_threshold(golden):
    """test_WorldFeatLabelGen.py:765-836: sum((world_amp - mcep80 -> amp)^2) < 100."""
    from idiaptts_b200.AudioProcessing import AudioProcessing
    from idiaptts_b200.compat import pysptk as ps
    x, c, f0, fs = golden_utterance(golden, "LJ001-0008")
    sp = world_np.cheaptrick(x, f0, world_np.temporal_positions(len(f0)), fs)
    amp = np.sqrt(sp)
    mc = AudioProcessing.extract_mcep(amp, 80, 0.41)
    assert mc.dtype == np.float32 and mc.shape == (len(f0), 80)
    rec = AudioProcessing.mcep_to_amp_sp(mc, fs, alpha=0.41)
    assert ((amp - rec) ** 2).sum() < 100
    ref = glue_np.mcep_to_amp_sp(mc, fs, alpha=0.41)
    assert (np.abs(rec - ref) / ref).max() < 5e-5
    mc60 = mc[:, :60].astype(np.float64)
    np.testing.assert_allclose(ps.mgc2sp(mc60, 0.41, 0.0, 1024).real, sptk_np.mgc2sp(mc60, 0.41, 0.0, 1024).real, atol=2e-5)
    p = ps.mc2sp(mc60, 0.41, 1024)
    assert (np.abs(p - sptk_np.mc2sp(mc60, 0.41, 1024)) / p).max() < 1e-4


@pytest.mark.parametrize("order,fft_size,nframes", [(59, 1024, 1000), (59, 1024, 130), (24, 512, 77), (39, 2048, 300)])
def test_tensor_core_mc2sp_vs_oracle_and_cuda_core_kernel(dev, order, fft_size, nframes):
    """The tcgen05 mel-cepstrum -> spectrum kernel of the batched synthesis path (float32 plane, order <= 59) against the fp64 oracle
    and against the CUDA-core kernel, for the three output forms (log amplitude, amplitude, float32 amplitude squared), float32 and
    float64 coefficients, a padded row stride, and frame counts that are not multiples of the 128-frame tile."""
    from idiaptts_b200 import ops
    rng = np.random.default_rng(order + nframes)
    mc = rng.standard_normal((nframes, order + 1)) * (0.6 ** np.arange(order + 1))[None, :]
    mc[:, 0] -= 4.0
    alpha = 0.455
    ref = sptk_np.mgc2sp(mc, alpha, 0.0, fft_size).real                       # log amplitude
    for dt in (torch.float32, torch.float64):
        wide = torch.zeros((nframes, order + 5), dtype=dt, device=dev)         # row stride > order + 1
        wide[:, :order + 1] = torch.from_numpy(mc).to(dev).to(dt)
        for do_exp, square in ((False, False), (True, False), (True, True)):
            kw = dict(scale=1.0, do_exp=do_exp, out_dtype=torch.float32, order=order, mc_stride=order + 5, square=square)
            y_tc = ops.mc2sp(wide, alpha, fft_size, impl="tc", **kw).cpu().numpy().astype(np.float64)
            y_cc = ops.mc2sp(wide, alpha, fft_size, impl="cc", **kw).cpu().numpy().astype(np.float64)
            want = ref if not do_exp else (np.exp(ref) ** 2 if square else np.exp(ref))
            assert np.isfinite(y_tc).all()
            if do_exp:
                assert (np.abs(y_tc - want) / want).max() < 5e-5 and (np.abs(y_tc - y_cc) / want).max() < 5e-5
            else:
                assert np.abs(y_tc - want).max() < 2e-5 and np.abs(y_tc - y_cc).max() < 2e-5


def test_ragged_batch_equals_single_utterances_and_edge_cases(golden, dev):
    from idiaptts_b200 import ops, pipeline
    ids = ["LJ001-0002", "LJ001-0008", "LJ001-0004"]
    waves = [golden[i + "/wav"] for i in ids]
    f0s = [np.where(golden[i + "/cmp"][:, 63] > 0, np.exp(golden[i + "/cmp"][:, 60].astype(np.float64)), 0.0) for i in ids]
    f0s[2] = np.zeros_like(f0s[2])       # an utterance without any voiced frame
    f0s[1][10:14] = 20.0                 # below CheapTrick's floor -> treated with the 500 Hz default window
    an = pipeline.WorldAnalyzer(16000, 60, 0.58, device=dev, chunk_frames=500)  # several chunks, chunk edge inside an utterance
    batch = ops.RaggedBatch.from_host(waves, f0s, 16000, device=dev)
    feats, sums, status = an.extract(batch)
    ops.raise_for_status(status, "ragged")
    feats = feats.cpu().numpy()
    assert feats.shape == (sum(len(f) for f in f0s), 63) and np.isfinite(feats).all()
    off = np.concatenate(([0], np.cumsum([len(f) for f in f0s])))
    for u in range(3):
        single, _, _ = pipeline.WorldAnalyzer(16000, 60, 0.58, device=dev).extract(
            ops.RaggedBatch.from_host([waves[u]], [f0s[u]], 16000, device=dev))
        assert np.array_equal(single.cpu().numpy(), feats[off[u]:off[u + 1]])  # batching does not change a single bit
    assert np.all(feats[off[2]:off[3], 61] == 0) and np.all(feats[off[2]:off[3], 60] == 0)  # no voiced frame: lf0 stays 0
    assert np.all(feats[off[2]:off[3], 62] == np.float32(-8.685697e-12))
    s = sums.cpu().numpy()
    np.testing.assert_allclose(s[:63], feats.astype(np.float64).sum(0), rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(s[63:], (feats.astype(np.float64) ** 2).sum(0), rtol=1e-9, atol=1e-6)
    # empty batch: no launch, empty outputs
    eb = ops.RaggedBatch.from_host([np.zeros(0, np.int16)], [np.zeros(0)], 16000, device=dev)
    ef, es, _ = an.extract(eb)
    assert ef.shape == (0, 63) and float(es.abs().sum()) == 0.0
    # an F0 far above the analysis range is reported, not silently mangled
    bad = ops.RaggedBatch.from_host([waves[0]], [np.full(len(f0s[0]), 7000.0)], 16000, device=dev)
    _, st = ops.cheaptrick(bad)
    with pytest.raises(ValueError, match="F0"):
        ops.raise_for_status(st, "cheaptrick")


def test_22k_and_48k_sizes(dev):
    """No reference goldens exist at these rates (smoke only in the reference, test_WorldFeatLabelGen.py:611-629): oracle parity."""
    from idiaptts_b200 import ops, synthetic
    for fs, n_fft, nap in ((22050, 1024, 2), (48000, 2048, 5)):
        waves, f0s = synthetic.make_corpus(1, fs, seed=3, mean_dur=0.6)
        w, f0 = waves[0].numpy(), f0s[0]
        batch = ops.RaggedBatch.from_host([w], [f0], fs, device=dev)
        sp, st = ops.cheaptrick(batch, out_dtype=torch.float64)
        assert sp.shape[1] == n_fft // 2 + 1
        x = w.astype(np.float64) / 32768.0
        t = world_np.temporal_positions(len(f0))
        ref = world_np.cheaptrick(x, f0, t, fs)
        assert (np.abs(sp.cpu().numpy() - ref) / ref).max() < 1e-6
        v_ref, c_ref = world_np.d4c_coarse(x, f0, t, fs)
        for precision, tol in (("f64", 1e-7), ("fast", 5e-4)):   # dB; the fast path runs its FFTs in single precision
            coarse, voiced, st = ops.d4c_coarse(batch, status=st, precision=precision)
            assert coarse.shape[1] == nap and np.array_equal(voiced.cpu().numpy().astype(bool), v_ref)  # decisions: bit-exact
            assert np.abs(coarse.cpu().numpy()[v_ref] - c_ref[v_ref]).max() < tol, precision
            bap = ops.bap_from_coarse(coarse, voiced, fs, n_fft).cpu().numpy()
            np.testing.assert_allclose(bap, world_np.code_aperiodicity(world_np.d4c(x, f0, t, fs), fs), atol=5e-4)
        assert ops.raise_for_status(st, "sizes") == 0


def test_padded_planes_give_identical_results(golden, dev):
    """The fused path pads envelope rows to a multiple of 8 floats (aligned 16-byte loads in the mel-cepstrum kernel): the row
    stride of b2w_cheaptrick / b2w_mcep_tc must not change a single bit of the results."""
    from idiaptts_b200 import ops
    x, c, f0, fs = golden_utterance(golden, "LJ001-0003")
    batch = ops.RaggedBatch.from_host([x], [f0], fs, device=dev, preemphasis=0.0)
    K = 513
    tight = torch.empty((len(f0), K), dtype=torch.float32, device=dev)
    padded = torch.full((len(f0), 520), float("nan"), dtype=torch.float32, device=dev)
    ops.cheaptrick(batch, fft_size=1024, out=tight)
    ops.cheaptrick(batch, fft_size=1024, out=padded[:, :K])
    assert torch.equal(tight, padded[:, :K]) and torch.isnan(padded[:, K:]).all()   # the padding is never written
    it_a = torch.zeros(len(f0), dtype=torch.int32, device=dev)
    it_b = torch.zeros_like(it_a)
    mc_a, st_a = ops.mcep(tight, 59, 0.58, is_power=True, iters=it_a)
    mc_b, st_b = ops.mcep(padded[:, :K], 59, 0.58, is_power=True, iters=it_b)                  # NaN padding must be ignored
    assert int(st_a.item()) == 0 and int(st_b.item()) == 0
    assert torch.equal(mc_a, mc_b) and torch.equal(it_a, it_b)
    mc_c, _ = ops.mcep(tight.double(), 59, 0.58, is_power=True)                                 # float64 plane: scalar-load path
    assert (mc_c - mc_a).abs().max().item() < 1e-5


def test_digital_silence_frames(dev):
    """WORLD adds two safeguard noises (randn() * 1e-12 to every windowed sample, eps * |randn()| to the smoothed spectrum); the
    oracle and the kernels use 0 and + eps instead.  On digital silence those terms are the ONLY signal: check
    that the kernels and the oracle agree there (they must: same substitution), that the envelope is the floor the substitution
    implies, that nothing is NaN, and that a frame half in silence is still within tolerance."""
    from idiaptts_b200 import ops
    fs = 16000
    x = np.zeros(8000)
    rng = np.random.default_rng(5)
    n = np.arange(4000)                             # the second half: a 150 Hz harmonic signal in a little noise
    x[4000:] = 0.01 * rng.standard_normal(4000) + sum(0.1 / h * np.sin(2 * np.pi * 150 * h * n / fs) for h in range(1, 20))
    f0 = np.zeros(101)
    f0[20:40] = 120.0                               # a "voiced" stretch inside the silence (what a bad F0 cache would give)
    f0[60:90] = 150.0
    t = world_np.temporal_positions(len(f0))
    ref = world_np.cheaptrick(x, f0, t, fs)
    batch = ops.RaggedBatch.from_host([x], [f0], fs, device=dev)
    for dt, tol in ((torch.float64, 1e-6), (torch.float32, 1e-4)):
        sp, st = ops.cheaptrick(batch, out_dtype=dt)
        out = sp.cpu().numpy().astype(np.float64)
        assert np.isfinite(out).all() and (out > 0).all()
        assert (np.abs(out - ref) / ref).max() < tol, dt
    silent = ref[:30]
    assert silent.max() < 1e-12                      # eps-floor envelope: exp(lifter(log(0 + eps))) ~ 2.2e-16
    mc, st = ops.mcep(torch.from_numpy(ref).to(dev), 59, 0.41, is_power=True)
    assert torch.isfinite(mc).all()
    # D4C: LoveTrain divides two zero band powers on silent "voiced" frames (NaN in WORLD too); decisions equal the oracle's
    v_ref, c_ref = world_np.d4c_coarse(x, f0, t, fs)
    for precision in ("f64", "fast"):
        coarse, voiced, _ = ops.d4c_coarse(batch, precision=precision)
        v = voiced.cpu().numpy().astype(bool)
        assert np.array_equal(v[60:90], v_ref[60:90])          # frames with signal: identical decisions
        m = v_ref & v
        m[:50] = False
        assert m.sum() >= 25                                    # LoveTrain keeps the harmonic stretch voiced
        assert np.abs(coarse.cpu().numpy()[m] - c_ref[m]).max() < (1e-7 if precision == "f64" else 5e-4)


def test_config2_shaped_utterances_vs_c_oracle(dev):
    """BASELINE.json configs[1] shape (6.5 s utterances at 22.05 kHz, alpha = mcepalpha(22050), nap = 2): the fused extraction and
    the batched synthesis against the C oracle on the SAME corpus generator bench.py uses -- the test-suite twin of the bench's
    parity block."""
    from idiaptts_b200 import ops, pipeline, synthetic
    from oracle import world_c
    fs = 22050
    waves, f0s = synthetic.make_corpus(3, fs, seed=2, mean_dur=6.5, std_dur=1.8, dur_quantum=0.1)
    an = pipeline.WorldAnalyzer(fs, 60, device=dev)
    batch = ops.RaggedBatch.from_host([w.numpy() for w in waves], f0s, fs, device=dev)
    feats, _, st = an.extract(batch)
    assert ops.raise_for_status(st, "extract") & ~8 == 0
    syn = pipeline.WorldSynthesizer(fs, 60, device=dev)
    y, out_off, st = syn.synthesize(feats, batch.frame_off)
    assert ops.raise_for_status(st, "synth") == 0
    fh, yh = feats.cpu().numpy(), y.cpu().numpy().astype(np.float64)
    fo = batch.frame_off.cpu().numpy()
    for u in range(3):
        ref = world_c.extract(waves[u].numpy(), fs, f0s[u], 60, an.alpha)
        g = fh[fo[u]:fo[u + 1]]
        assert g.shape == ref.shape == (world_np.num_frames(len(waves[u]), fs), 64)      # frame count: bit-exact
        assert glue_np.mcd_db(ref[:, :60], g[:, :60]) < 0.01
        assert np.array_equal(ref[:, 61], g[:, 61])                                      # vuv: bit-exact
        assert np.abs(ref[:, 60] - g[:, 60]).max() < 1e-5 and np.abs(ref[:, 62:] - g[:, 62:]).max() < 1e-3
        y_ref = world_c.synthesize_features(g, fs, 60, an.alpha).astype(np.float64)
        got = yh[out_off[u]:out_off[u + 1]]
        assert len(got) == len(y_ref)
        assert 10 * np.log10((y_ref ** 2).sum() / ((got - y_ref) ** 2).sum()) > 60
