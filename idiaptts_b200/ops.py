"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out, every computation done by
libb200world.so through the C ABI (include/b200world.h).  torch is plumbing here (memory, streams), not compute.

There is no CPU path: CPU tensors are rejected."""
import math
import threading

import numpy as np
import torch

from . import _lib
from ._lib import B2W_F32, B2W_F64, B2W_I16, check

_DT = {torch.float64: B2W_F64, torch.float32: B2W_F32, torch.int16: B2W_I16}


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("idiaptts_b200 operators need CUDA tensors (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("all tensors must live on the same device")
        if not t.is_contiguous():
            raise ValueError("tensors must be contiguous")
    return dev


# ----------------------------------------------------------------------------------------------------------------------
# scalar helpers (pure functions of the sampling rate; computed by the library so that both sides agree)
# ----------------------------------------------------------------------------------------------------------------------
def get_cheaptrick_fft_size(fs, f0_floor=71.0):
    return int(_lib.load().b2w_cheaptrick_fft_size(int(fs), float(f0_floor)))


def get_num_aperiodicities(fs):
    return int(_lib.load().b2w_num_aperiodicities(int(fs)))


def get_d4c_fft_size(fs):
    return int(_lib.load().b2w_d4c_fft_size(int(fs)))


def num_frames(num_samples, fs, frame_period=5.0):
    """Frame count of pyworld.wav2world / dio (WORLD GetSamplesForDIO)."""
    return int(1000.0 * num_samples / fs / frame_period) + 1


# ----------------------------------------------------------------------------------------------------------------------
# ragged batch
# ----------------------------------------------------------------------------------------------------------------------
class RaggedBatch:
    """A set of utterances packed for the analysis kernels: waveform samples, cached F0 track, temporal positions.

    x           packed samples (float64 | float32 | int16, int16 means value / 32768)
    sample_off  int64 [U + 1]      frame_off  int64 [U + 1]
    f0, t       float64 [F]        frame_utt  int32 [F]
    """

    def __init__(self, x, sample_off, f0, t, frame_off, frame_utt, fs, preemphasis=0.0):
        self.device = _need_cuda(x, sample_off, f0, t, frame_off, frame_utt)
        if x.dtype not in _DT:
            raise ValueError("waveform dtype %s not supported" % x.dtype)
        assert sample_off.dtype == torch.int64 and frame_off.dtype == torch.int64 and frame_utt.dtype == torch.int32
        assert f0.dtype == torch.float64 and t.dtype == torch.float64
        self.x, self.sample_off, self.f0, self.t = x, sample_off, f0, t
        self.frame_off, self.frame_utt = frame_off, frame_utt
        self.fs = int(fs)
        self.preemphasis = float(preemphasis)
        self.num_utts = sample_off.numel() - 1
        self.num_frames = f0.numel()

    @staticmethod
    def from_host(waves, f0s, fs, frame_period=5.0, preemphasis=0.0, device="cuda", ts=None, pin=True):
        """waves: list of 1-D numpy arrays (float64/float32/int16); f0s: list of float64 F0 tracks (Hz, 0 = unvoiced),
        one value per frame.  ts: optional list of temporal positions; default i * frame_period / 1000."""
        device = torch.device(device)
        dt = waves[0].dtype
        lens = np.array([len(w) for w in waves], np.int64)
        flens = np.array([len(f) for f in f0s], np.int64)
        sample_off = np.concatenate(([0], np.cumsum(lens)))
        frame_off = np.concatenate(([0], np.cumsum(flens)))
        x = np.concatenate([np.ascontiguousarray(w, dt) for w in waves]) if len(waves) > 1 else np.ascontiguousarray(waves[0])
        f0 = np.concatenate([np.asarray(f, np.float64) for f in f0s])
        if ts is None:
            t = np.concatenate([np.arange(n) * frame_period / 1000.0 for n in flens])
        else:
            t = np.concatenate([np.asarray(v, np.float64) for v in ts])
        frame_utt = np.repeat(np.arange(len(waves), dtype=np.int32), flens)

        def up(a):
            h = torch.from_numpy(np.ascontiguousarray(a))
            if pin and h.numel() > 0:
                h = h.pin_memory()
            return h.to(device, non_blocking=True)

        return RaggedBatch(up(x), up(sample_off), up(f0), up(t), up(frame_off), up(frame_utt), fs, preemphasis)

    @staticmethod
    def from_packed(samples, sample_off, f0s, fs, frame_period=5.0, preemphasis=0.0, device="cuda"):
        """The same batch from an already packed host waveform (corpus_io.read_wavs_i16: one pinned int16 tensor + offsets):
        no per-utterance concatenation on the host.  f0s: list of float64 tracks, one value per frame."""
        device = torch.device(device)
        flens = np.array([len(f) for f in f0s], np.int64)
        frame_off = np.concatenate(([0], np.cumsum(flens)))
        f0 = np.concatenate([np.asarray(f, np.float64) for f in f0s]) if len(f0s) else np.zeros(0)
        frame_utt = np.repeat(np.arange(len(f0s), dtype=np.int32), flens)
        # i * frame_period / 1000 per utterance, on the device: the same two IEEE operations numpy does in from_host
        fo_dev = torch.from_numpy(frame_off).to(device)
        fu_dev = torch.from_numpy(frame_utt).to(device)
        idx = torch.arange(int(frame_off[-1]), device=device, dtype=torch.int64) - fo_dev[:-1][fu_dev.long()]
        t = idx.to(torch.float64) * frame_period / 1000.0

        def up(a):
            h = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
            return h.to(device, non_blocking=h.is_pinned())

        return RaggedBatch(up(samples), up(np.asarray(sample_off, np.int64)), up(f0), t, fo_dev, fu_dev, fs, preemphasis)

    def c_struct(self, frame_lo=0, frame_hi=None):
        """struct b2w_batch for frames [frame_lo, frame_hi) (chunking keeps the intermediate planes bounded)."""
        if frame_hi is None:
            frame_hi = self.num_frames
        b = _lib.Batch()
        b.x = self.x.data_ptr()
        b.x_dtype = _DT[self.x.dtype]
        b.num_utts = self.num_utts
        b.preemphasis = self.preemphasis
        b.utt_sample_offset = self.sample_off.data_ptr()
        b.frame_utt = self.frame_utt.data_ptr() + 4 * frame_lo
        b.f0 = self.f0.data_ptr() + 8 * frame_lo
        b.t = self.t.data_ptr() + 8 * frame_lo
        b.num_frames = frame_hi - frame_lo
        b.fs = self.fs
        return b


def _timed(events, name, units, fn):
    """Runs fn(); with an `events` list also brackets it with CUDA events on the current stream and appends
    (name, units, start, end) -- bench.py's per-kernel timing."""
    if events is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    events.append((name, units, e0, e1))
    return r


def new_status(device):
    return torch.zeros(1, dtype=torch.int32, device=device)


def raise_for_status(status, what):
    """Reads the device status word (synchronises) and raises like pyworld / pysptk would."""
    s = int(status.item())
    if s & _lib.STATUS_ZERO_PERIODOGRAM:
        raise RuntimeError("%s: zero(s) are found in periodogram, use eps option to floor" % what)
    if s & _lib.STATUS_SOLVE_FAILED:
        raise RuntimeError("%s: failed to compute mcep; error occured in theq" % what)
    if s & _lib.STATUS_F0_TOO_HIGH:
        raise ValueError("%s: an F0 value is too high for the analysis window / pulse buffer" % what)
    return s


# ----------------------------------------------------------------------------------------------------------------------
# F0 estimation: DIO + StoneMask
# ----------------------------------------------------------------------------------------------------------------------
DIO_MAX_CHUNK_SAMPLES = 96 * 1024 * 1024  # the event lists cost ~120 bytes of workspace per sample: 96 M samples ~ 12 GB (of 180 GB HBM)


def _utt_chunks(sample_off_host, max_samples):
    """Utterance ranges [u0, u1) whose sample counts stay below max_samples (a single longer utterance forms its own chunk)."""
    U = len(sample_off_host) - 1
    u0 = 0
    while u0 < U:
        u1 = u0 + 1
        while u1 < U and sample_off_host[u1 + 1] - sample_off_host[u0] <= max_samples:
            u1 += 1
        yield u0, u1
        u0 = u1


def _sub_batch(batch, u0, u1, s_off, f_off, f0):
    """struct b2w_batch + frame offsets of utterances [u0, u1) (offsets rebased; the tensors are kept alive by the caller)."""
    so = (batch.sample_off[u0:u1 + 1] - int(s_off[u0])).contiguous()
    fo = (batch.frame_off[u0:u1 + 1] - int(f_off[u0])).contiguous()
    fu = (batch.frame_utt[int(f_off[u0]):int(f_off[u1])] - u0).contiguous()
    b = _lib.Batch()
    b.x = batch.x.data_ptr() + int(s_off[u0]) * batch.x.element_size()
    b.x_dtype = _DT[batch.x.dtype]
    b.num_utts = u1 - u0
    b.preemphasis = batch.preemphasis
    b.utt_sample_offset = so.data_ptr()
    b.frame_utt = fu.data_ptr()
    b.f0 = 0 if f0 is None else f0.data_ptr() + 8 * int(f_off[u0])
    b.t = batch.t.data_ptr() + 8 * int(f_off[u0])
    b.num_frames = int(f_off[u1] - f_off[u0])
    b.fs = batch.fs
    return b, (so, fo, fu)


def dio(batch, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2.0, frame_period=5.0, allowed_range=0.1, step2="erosion",
        max_chunk_samples=None):
    """pyworld.dio (speed = 1) on a ragged batch -> f0 [F] float64 (batch.f0 is ignored; batch.t must be i * frame_period / 1000).

    step2: "erosion" reproduces the reference's fixtures (oracle/dio_np.py), "sections" is the later WORLD variant."""
    if step2 not in ("erosion", "sections"):
        raise ValueError("step2 must be 'erosion' or 'sections'")
    lib = _lib.load()
    dev = batch.device
    if lib.b2w_dio_num_bands(float(f0_floor), float(f0_ceil), float(channels_in_octave)) < 1:
        raise ValueError("bad f0_floor / f0_ceil / channels_in_octave")
    out = torch.empty(batch.num_frames, dtype=torch.float64, device=dev)
    if batch.num_frames == 0:
        return out
    s_off = batch.sample_off.cpu().numpy()
    f_off = batch.frame_off.cpu().numpy()
    with torch.cuda.device(dev):
        for u0, u1 in _utt_chunks(s_off, max_chunk_samples or DIO_MAX_CHUNK_SAMPLES):
            S = int(s_off[u1] - s_off[u0])
            F = int(f_off[u1] - f_off[u0])
            if F == 0:
                continue
            nbytes = int(lib.b2w_dio_workspace_bytes(S, u1 - u0, F, batch.fs, float(f0_floor), float(f0_ceil), float(channels_in_octave)))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            b, keep = _sub_batch(batch, u0, u1, s_off, f_off, None)
            check(lib.b2w_dio(b, S, keep[1].data_ptr(), float(f0_floor), float(f0_ceil), float(channels_in_octave), float(frame_period),
                              float(allowed_range), 1 if step2 == "sections" else 0, ws.data_ptr(),
                              out.data_ptr() + 8 * int(f_off[u0]), _stream(dev)), "b2w_dio")
            del ws, keep  # allocated and used on the current stream: the caching allocator's stream-ordered reuse is safe
    return out


def dio_fir_taps(fs, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2.0):
    """Taps per sample of DIO's two FIR stages (low cut + all bands): the kernel's algorithmic DFMA count per sample."""
    nb = 1 + int(math.log(f0_ceil / f0_floor) / math.log(2.0) * channels_in_octave)
    rnd = lambda v: int(v + 0.5)
    taps = 2 * rnd(fs / 50.0) + 1
    for i in range(nb):
        taps += 4 * rnd(fs / (f0_floor * 2.0 ** ((i + 1) / channels_in_octave)) / 2.0)
    return taps


def stonemask(batch, f0=None):
    """pyworld.stonemask on a ragged batch: refines f0 (default batch.f0) -> [F] float64."""
    lib = _lib.load()
    dev = batch.device
    f0 = batch.f0 if f0 is None else f0
    _need_cuda(f0)
    assert f0.dtype == torch.float64 and f0.numel() == batch.num_frames
    out = torch.empty(batch.num_frames, dtype=torch.float64, device=dev)
    if batch.num_frames == 0:
        return out
    b = batch.c_struct()
    b.f0 = f0.data_ptr()
    with torch.cuda.device(dev):
        check(lib.b2w_stonemask(b, out.data_ptr(), _stream(dev)), "b2w_stonemask")
    return out


def estimate_f0(batch, frame_period=5.0, **dio_args):
    """The F0 half of pyworld.wav2world (WorldFeatLabelGen.py:792): dio + stonemask; also stores the result in batch.f0."""
    f0 = stonemask(batch, dio(batch, frame_period=frame_period, **dio_args))
    batch.f0 = f0
    return f0


# ----------------------------------------------------------------------------------------------------------------------
# analysis
# ----------------------------------------------------------------------------------------------------------------------
def cheaptrick(batch, fft_size=None, q1=-0.15, out_dtype=torch.float64, status=None, frame_lo=0, frame_hi=None, out=None):
    """pyworld.cheaptrick on a ragged batch -> sp [F, fft_size/2+1] (power)."""
    lib = _lib.load()
    if fft_size is None:
        fft_size = get_cheaptrick_fft_size(batch.fs)
    if frame_hi is None:
        frame_hi = batch.num_frames
    nf = frame_hi - frame_lo
    K = fft_size // 2 + 1
    if out is None:
        out = torch.empty((nf, K), dtype=out_dtype, device=batch.device)
    if status is None:
        status = new_status(batch.device)
    b = batch.c_struct(frame_lo, frame_hi)
    with torch.cuda.device(batch.device):
        assert out.dim() == 2 and out.shape[1] == K and out.stride(1) == 1 and out.shape[0] >= nf
        check(lib.b2w_cheaptrick(b, fft_size, float(q1), out.data_ptr(), _DT[out.dtype], int(out.stride(0)), status.data_ptr(),
                                 _stream(batch.device)), "b2w_cheaptrick")
    return out, status


def d4c_coarse(batch, threshold=0.85, status=None, frame_lo=0, frame_hi=None, precision="fast"):
    """LoveTrain + D4C band aperiodicity -> (coarse_db [F, nap] f64, voiced [F] uint8).
    precision: "fast" = single-precision FFTs with fp64 cumulative sums and an fp64 re-evaluation of the frames whose LoveTrain
    ratio is too close to the threshold to call (b2w_d4c_coarse: decisions identical to "f64", coarse_db within ~1e-4 dB);
    "f64" = double precision throughout (b2w_d4c_coarse_f64)."""
    assert precision in ("fast", "f64")
    lib = _lib.load()
    if frame_hi is None:
        frame_hi = batch.num_frames
    nf = frame_hi - frame_lo
    nap = get_num_aperiodicities(batch.fs)
    coarse = torch.zeros((nf, max(nap, 1)), dtype=torch.float64, device=batch.device)
    voiced = torch.zeros((nf,), dtype=torch.uint8, device=batch.device)
    if status is None:
        status = new_status(batch.device)
    b = batch.c_struct(frame_lo, frame_hi)
    with torch.cuda.device(batch.device):
        fn = lib.b2w_d4c_coarse if precision == "fast" else lib.b2w_d4c_coarse_f64
        check(fn(b, float(threshold), coarse.data_ptr(), voiced.data_ptr(), status.data_ptr(), _stream(batch.device)),
              "b2w_d4c_coarse")
    return coarse, voiced, status


def d4c_expand(coarse, voiced, fs, fft_size):
    lib = _lib.load()
    dev = _need_cuda(coarse, voiced)
    F = voiced.numel()
    ap = torch.empty((F, fft_size // 2 + 1), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.b2w_d4c_expand(coarse.data_ptr(), voiced.data_ptr(), F, int(fs), int(fft_size), ap.data_ptr(), _stream(dev)),
              "b2w_d4c_expand")
    return ap


def bap_from_coarse(coarse, voiced, fs, fft_size, out=None, out_stride=None):
    """code_aperiodicity(d4c(...)) without the [F, K] plane.  out: float32 tensor written with row stride out_stride."""
    lib = _lib.load()
    dev = _need_cuda(coarse, voiced)
    F = voiced.numel()
    nap = get_num_aperiodicities(fs)
    if out is None:
        out = torch.empty((F, nap), dtype=torch.float32, device=dev)
        out_stride = nap
    with torch.cuda.device(dev):
        check(lib.b2w_bap_from_coarse(coarse.data_ptr(), voiced.data_ptr(), F, int(fs), int(fft_size), out.data_ptr(),
                                      int(out_stride), _stream(dev)), "b2w_bap_from_coarse")
    return out


def code_aperiodicity(ap, fs):
    lib = _lib.load()
    dev = _need_cuda(ap)
    assert ap.dtype == torch.float64 and ap.dim() == 2
    F, K = ap.shape
    nap = get_num_aperiodicities(fs)
    out = torch.empty((F, nap), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.b2w_code_aperiodicity(ap.data_ptr(), F, int(fs), 2 * (K - 1), out.data_ptr(), _stream(dev)),
              "b2w_code_aperiodicity")
    return out


def decode_aperiodicity(bap, fs, fft_size, out=None, out_dtype=torch.float64):
    """pyworld.decode_aperiodicity on [F, nap] float64 -> [F, K] float64 (or float32 for the fast synthesis path)."""
    lib = _lib.load()
    dev = _need_cuda(bap)
    assert bap.dtype == torch.float64 and bap.dim() == 2
    nap = get_num_aperiodicities(fs)
    if bap.shape[1] != nap:
        raise ValueError("coded aperiodicity has %d bands, fs=%d needs %d" % (bap.shape[1], fs, nap))
    F = bap.shape[0]
    if out is None:
        out = torch.empty((F, fft_size // 2 + 1), dtype=out_dtype, device=dev)
    assert out.dtype in (torch.float64, torch.float32) and out.shape == (F, fft_size // 2 + 1) and out.is_contiguous()
    fn = lib.b2w_decode_aperiodicity if out.dtype == torch.float64 else lib.b2w_decode_aperiodicity_f32
    with torch.cuda.device(dev):
        check(fn(bap.data_ptr(), F, int(fs), int(fft_size), out.data_ptr(), _stream(dev)), "b2w_decode_aperiodicity")
    return out


# ----------------------------------------------------------------------------------------------------------------------
# mel-cepstrum
# ----------------------------------------------------------------------------------------------------------------------
class McepTables:
    """The three precomputed all-pass warping matrices of (order, alpha, fft_size), float32 on one device."""
    _cache = {}
    _lock = threading.Lock()

    def __init__(self, order, alpha, fft_size, device):
        lib = _lib.load()
        K = fft_size // 2 + 1
        self.order, self.alpha, self.fft_size = int(order), float(alpha), int(fft_size)
        np0, np2, mp = lib.b2w_mcep_pad(order + 2), lib.b2w_mcep_pad(2 * order + 1), lib.b2w_mcep_pad(order + 1)
        m0t = np.empty((K, np0), np.float64)
        cmat = np.empty((mp, K), np.float64)
        m2t = np.empty((K, np2), np.float64)
        check(lib.b2w_mcep_tables_host(self.order, self.alpha, self.fft_size, m0t.ctypes.data, cmat.ctypes.data,
                                       m2t.ctypes.data), "b2w_mcep_tables_host")
        self.host64 = (m0t, cmat, m2t)
        self.m0t = torch.from_numpy(m0t.astype(np.float32)).to(device)
        self.cmat = torch.from_numpy(cmat.astype(np.float32)).to(device)
        self.m2t = torch.from_numpy(m2t.astype(np.float32)).to(device)
        # pre-tiled hi/lo TF32 streams of the tensor-core kernel (order <= 59: the solver workspaces of larger orders do not fit)
        self.stream0 = self.stream1 = None
        if self.order <= 59:
            n = int(lib.b2w_mcep_tc_stream_floats(self.fft_size))
            self.stream0 = torch.empty(n, dtype=torch.float32, device=device)
            self.stream1 = torch.empty(n, dtype=torch.float32, device=device)
            with torch.cuda.device(device):
                check(lib.b2w_mcep_tc_pretile(self.order, self.fft_size, self.m0t.data_ptr(), self.cmat.data_ptr(), self.m2t.data_ptr(),
                                              self.stream0.data_ptr(), self.stream1.data_ptr(), _stream(torch.device(device))),
                      "b2w_mcep_tc_pretile")
            torch.cuda.current_stream(torch.device(device)).synchronize()  # tables are shared by every stream afterwards

    @classmethod
    def get(cls, order, alpha, fft_size, device):
        key = (int(order), float(alpha), int(fft_size), str(torch.device(device)))
        with cls._lock:
            tab = cls._cache.get(key)
            if tab is None:
                tab = cls(order, alpha, fft_size, device)
                cls._cache[key] = tab
            return tab


def mcep(plane, order, alpha, is_power=False, miniter=2, maxiter=30, threshold=0.001, eps=1.0e-8, out=None,
         out_stride=None, out_dtype=torch.float32, iters=None, status=None, impl=None):
    """pysptk.mcep(itype=3 (amplitude) or 4 (power), etype=1) on a [F, K] plane -> mc [F, order+1].
    impl: "tc" = tcgen05 tensor-core kernel (default for order <= 59), "cc" = CUDA-core kernel (any order <= 127)."""
    lib = _lib.load()
    assert plane.dim() == 2 and plane.dtype in (torch.float32, torch.float64)
    F, K = plane.shape
    if plane.stride(1) != 1 or (F > 1 and plane.stride(0) < K):
        plane = plane.contiguous()
    in_stride = int(plane.stride(0)) if F > 1 else K  # rows may be padded (the stride of a single row is arbitrary in torch)
    if not plane.is_cuda:
        raise ValueError("idiaptts_b200 operators need CUDA tensors (there is no CPU fallback)")
    dev = plane.device
    fft_size = 2 * (K - 1)
    tab = McepTables.get(order, alpha, fft_size, dev)
    if out is None:
        out = torch.empty((F, order + 1), dtype=out_dtype, device=dev)
        out_stride = order + 1
    if status is None:
        status = new_status(dev)
    if impl is None:
        impl = "tc" if tab.stream0 is not None else "cc"
    if impl == "tc":
        if tab.stream0 is None:
            raise ValueError("the tensor-core mcep kernel supports order <= 59")
        with torch.cuda.device(dev):
            check(lib.b2w_mcep_tc(plane.data_ptr(), _DT[plane.dtype], 1 if is_power else 0, in_stride, F, fft_size,
                                  int(order), float(alpha),
                                  int(miniter), int(maxiter), float(threshold), float(eps), tab.stream0.data_ptr(),
                                  tab.stream1.data_ptr(), out.data_ptr(), _DT[out.dtype], int(out_stride), _ptr(iters),
                                  status.data_ptr(), _stream(dev)), "b2w_mcep_tc")
        return out, status
    if in_stride != K:
        plane = plane.contiguous()
    with torch.cuda.device(dev):
        check(lib.b2w_mcep(plane.data_ptr(), _DT[plane.dtype], 1 if is_power else 0, F, fft_size, int(order), float(alpha),
                           int(miniter), int(maxiter), float(threshold), float(eps), tab.m0t.data_ptr(), tab.cmat.data_ptr(),
                           tab.m2t.data_ptr(), out.data_ptr(), _DT[out.dtype], int(out_stride), _ptr(iters),
                           status.data_ptr(), _stream(dev)), "b2w_mcep")
    return out, status


class MgcTables:
    """cos / sin tables of the all-pass warped frequency for the generalised mel-cepstrum kernels (b2w_mgcep, b2w_mgc2sp):
    forward [pad4(order + 1), K] (coefficients -> spectrum) and reduction [K, pad4(2 order + 1)] with the bin weights of the
    mean over the circle folded in (per-bin weights -> gradient / Toeplitz / Hankel sequences); float64 on the host, float32
    on the device.  (order, alpha, fft_size) -> cached per device."""
    _cache = {}
    _lock = threading.Lock()

    def __init__(self, order, alpha, fft_size, device):
        lib = _lib.load()
        K = fft_size // 2 + 1
        mp, np2 = lib.b2w_mcep_pad(order + 1), lib.b2w_mcep_pad(2 * order + 1)
        w = 2.0 * np.pi * np.arange(K) / fft_size
        wt = w + 2.0 * np.arctan2(alpha * np.sin(w), 1.0 - alpha * np.cos(w))   # warped frequency of the first-order all-pass
        fwd_cos, fwd_sin = np.zeros((mp, K)), np.zeros((mp, K))
        k = np.arange(order + 1)[:, None]
        fwd_cos[:order + 1], fwd_sin[:order + 1] = np.cos(k * wt[None, :]), np.sin(k * wt[None, :])
        W = np.full(K, 2.0 / fft_size)
        W[0] = W[-1] = 1.0 / fft_size
        red_cos, red_sin = np.zeros((K, np2)), np.zeros((K, np2))
        n = np.arange(2 * order + 1)[None, :]
        red_cos[:, :2 * order + 1] = W[:, None] * np.cos(wt[:, None] * n)
        red_sin[:, :2 * order + 1] = W[:, None] * np.sin(wt[:, None] * n)
        self.host64 = (fwd_cos, fwd_sin, red_cos, red_sin)
        to = lambda a: torch.from_numpy(a.astype(np.float32)).to(device)
        self.fwd_cos, self.fwd_sin, self.red_cos, self.red_sin = to(fwd_cos), to(fwd_sin), to(red_cos), to(red_sin)

    @classmethod
    def get(cls, order, alpha, fft_size, device):
        key = (int(order), float(alpha), int(fft_size), str(torch.device(device)))
        with cls._lock:
            tab = cls._cache.get(key)
            if tab is None:
                tab = cls(order, alpha, fft_size, device)
                cls._cache[key] = tab
            return tab


def mgcep(plane, order, alpha, gamma, is_power=False, miniter=2, maxiter=30, threshold=0.001, eps=1.0e-8, out=None,
          out_stride=None, out_dtype=torch.float32, iters=None, status=None):
    """pysptk.mgcep(itype=3 (amplitude) or 4 (power), etype=1, otype=0) on a [F, K] plane -> mgc [F, order+1] (SURVEY 8f N3;
    parity unpinned, see csrc/mgcep.cu).  gamma in [-1, 0); gamma = 0 is `mcep`."""
    if gamma == 0.0:
        return mcep(plane, order, alpha, is_power, miniter, maxiter, threshold, eps, out, out_stride, out_dtype, iters, status)
    lib = _lib.load()
    assert plane.dim() == 2 and plane.dtype in (torch.float32, torch.float64)
    if not plane.is_cuda:
        raise ValueError("idiaptts_b200 operators need CUDA tensors (there is no CPU fallback)")
    plane = plane.contiguous()
    F, K = plane.shape
    dev = plane.device
    fft_size = 2 * (K - 1)
    tab0 = McepTables.get(order, alpha, fft_size, dev)
    tab = MgcTables.get(order, alpha, fft_size, dev)
    if out is None:
        out = torch.empty((F, order + 1), dtype=out_dtype, device=dev)
        out_stride = order + 1
    if status is None:
        status = new_status(dev)
    with torch.cuda.device(dev):
        check(lib.b2w_mgcep(plane.data_ptr(), _DT[plane.dtype], 1 if is_power else 0, F, fft_size, int(order), float(gamma),
                            int(miniter), int(maxiter), float(threshold), float(eps), tab0.m0t.data_ptr(), tab.fwd_cos.data_ptr(),
                            tab.fwd_sin.data_ptr(), tab.red_cos.data_ptr(), tab.red_sin.data_ptr(), out.data_ptr(), _DT[out.dtype],
                            int(out_stride), _ptr(iters), status.data_ptr(), _stream(dev)), "b2w_mgcep")
    return out, status


def mgc2sp(mgc, alpha, gamma, fft_size, out_dtype=torch.float32, order=None, mgc_stride=None):
    """Amplitude spectrum |H| [F, K] of generalised mel-cepstra (AudioProcessing.mgc_to_amp_sp = exp(Re pysptk.mgc2sp));
    gamma = 0 is mc2sp."""
    if gamma == 0.0:
        return mc2sp(mgc, alpha, fft_size, out_dtype=out_dtype, order=order, mc_stride=mgc_stride)
    lib = _lib.load()
    dev = _need_cuda(mgc)
    assert mgc.dim() == 2 and mgc.dtype in (torch.float32, torch.float64)
    F = mgc.shape[0]
    if order is None:
        order = mgc.shape[1] - 1
    if mgc_stride is None:
        mgc_stride = mgc.shape[1]
    tab = MgcTables.get(order, alpha, fft_size, dev)
    out = torch.empty((F, fft_size // 2 + 1), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        check(lib.b2w_mgc2sp(mgc.data_ptr(), _DT[mgc.dtype], int(mgc_stride), F, int(fft_size), int(order), float(gamma),
                             tab.fwd_cos.data_ptr(), tab.fwd_sin.data_ptr(), out.data_ptr(), _DT[out.dtype], _stream(dev)),
              "b2w_mgc2sp")
    return out


def merlin_post_filter(mgc, alpha, fft_size=1024, coef=1.4):
    """nnmnkwii.postfilters.merlin_post_filter on [F, D] mel-cepstra (device): coefficients from c(2) on are scaled by `coef`
    and c(0) is shifted so that the energy r(0) = mean |H|^2 of the minimum-phase response is unchanged (mc2b / b2mc only
    move c(0) here).  The two energies come from the mel-cepstrum -> power-spectrum kernel (b2w_mc2sp) and a weighted row sum."""
    dev = _need_cuda(mgc)
    assert mgc.dim() == 2
    x = mgc.to(torch.float32).contiguous()
    F, D = x.shape
    weight = torch.full((D,), float(coef), dtype=torch.float32, device=dev)
    weight[:2] = 1.0
    scaled = (x * weight).contiguous()
    K = fft_size // 2 + 1
    W = torch.full((K,), 2.0 / fft_size, dtype=torch.float64, device=dev)
    W[0] = W[-1] = 1.0 / fft_size
    r0 = mc2sp(x, alpha, fft_size, scale=2.0, do_exp=True, out_dtype=torch.float64) @ W
    p_r0 = mc2sp(scaled, alpha, fft_size, scale=2.0, do_exp=True, out_dtype=torch.float64) @ W
    out = scaled.double()
    out[:, 0] += 0.5 * torch.log(r0 / p_r0)
    return out


def mc2sp(mc, alpha, fft_size, scale=1.0, do_exp=True, out_dtype=torch.float32, order=None, mc_stride=None, square=False, out=None,
          impl=None):
    """(exp of) scale * Re FFT(freqt(mc, fft_size/2, -alpha)): log-amplitude / amplitude / power spectrum from mel-cepstra.
    square=True (with do_exp): the float32 amplitude squared in float64, i.e. world_features_to_raw's pow_sp.
    impl: "tc" = tcgen05 tensor-core kernel (default for a float32 plane and order <= 59), "cc" = CUDA-core kernel."""
    lib = _lib.load()
    dev = _need_cuda(mc)
    assert mc.dim() == 2 and mc.dtype in (torch.float32, torch.float64)
    F = mc.shape[0]
    if order is None:
        order = mc.shape[1] - 1
    if mc_stride is None:
        mc_stride = mc.shape[1]
    tab = McepTables.get(order, alpha, fft_size, dev)
    if out is None:
        out = torch.empty((F, fft_size // 2 + 1), dtype=out_dtype, device=dev)
    assert out.dtype == out_dtype and out.shape == (F, fft_size // 2 + 1) and out.is_contiguous()
    if impl is None:
        impl = "tc" if (out.dtype == torch.float32 and tab.stream1 is not None and order >= 1) else "cc"
    if impl == "tc":
        if out.dtype != torch.float32 or tab.stream1 is None:
            raise ValueError("the tensor-core mc2sp kernel writes float32 planes for order <= 59")
        with torch.cuda.device(dev):
            check(lib.b2w_mc2sp_tc(mc.data_ptr(), _DT[mc.dtype], int(mc_stride), F, int(fft_size), int(order), tab.stream1.data_ptr(),
                                   float(scale), (2 if square else 1) if do_exp else 0, out.data_ptr(), _stream(dev)), "b2w_mc2sp_tc")
        return out
    with torch.cuda.device(dev):
        check(lib.b2w_mc2sp(mc.data_ptr(), _DT[mc.dtype], int(mc_stride), F, int(fft_size), int(order), tab.cmat.data_ptr(),
                            float(scale), (2 if square else 1) if do_exp else 0, out.data_ptr(), _DT[out.dtype], _stream(dev)),
              "b2w_mc2sp")
    return out


# ----------------------------------------------------------------------------------------------------------------------
# labels, deltas, statistics
# ----------------------------------------------------------------------------------------------------------------------
def lf0_vuv(f0, frame_off, f0_silence_threshold=30, lf0_zero=0, lf0_out=None, vuv_out=None, out_stride=1):
    lib = _lib.load()
    dev = _need_cuda(f0, frame_off)
    F = f0.numel()
    if lf0_out is None:
        lf0_out = torch.empty((F, 1), dtype=torch.float32, device=dev)
        vuv_out = torch.empty((F, 1), dtype=torch.float32, device=dev)
        out_stride = 1
    with torch.cuda.device(dev):
        check(lib.b2w_lf0_vuv(f0.data_ptr(), frame_off.data_ptr(), frame_off.numel() - 1, float(f0_silence_threshold),
                              float(lf0_zero), lf0_out.data_ptr(), vuv_out.data_ptr(), int(out_stride), _stream(dev)),
              "b2w_lf0_vuv")
    return lf0_out, vuv_out


def deltas(feats, frame_off, want_double=True):
    """np.gradient deltas (and double deltas) per utterance of a [F, D] float32 matrix."""
    lib = _lib.load()
    dev = _need_cuda(feats, frame_off)
    assert feats.dtype == torch.float32 and feats.dim() == 2
    F, D = feats.shape
    d = torch.empty_like(feats)
    dd = torch.empty_like(feats) if want_double else None
    with torch.cuda.device(dev):
        check(lib.b2w_deltas(feats.data_ptr(), D, D, frame_off.data_ptr(), frame_off.numel() - 1, F, d.data_ptr(), _ptr(dd), D,
                             _stream(dev)), "b2w_deltas")
    return d, dd


def stats_accumulate(feats, sums, gram=None, dim=None, stride=None, num_frames=None, offset_elems=0):
    """sums [2*dim] f64 += (sum x, sum x^2); gram [dim, dim] f64 += X^T X."""
    lib = _lib.load()
    dev = _need_cuda(feats, sums, gram)
    assert feats.dtype == torch.float32 and sums.dtype == torch.float64
    if dim is None:
        num_frames, dim = feats.shape
        stride = dim
    with torch.cuda.device(dev):
        check(lib.b2w_stats_accumulate(feats.data_ptr() + 4 * offset_elems, int(stride), int(dim), int(num_frames),
                                       sums.data_ptr(), _ptr(gram), _stream(dev)), "b2w_stats_accumulate")
    return sums


# ----------------------------------------------------------------------------------------------------------------------
# trainer-facing batch
# ----------------------------------------------------------------------------------------------------------------------
def pad_normalise(feats, frame_off, mean=None, std_dev=None, batch_first=False, min_frames=None, want_mask=True, lengths=None):
    """Ragged rows [F, W] f32 -> (padded [T_max, U, W] (or [U, T_max, W]), mask [T_max, U, 1] (or [U, T_max, 1]) | None,
    lengths int64 [U] (host numpy)): preprocess_sample + prepare_batch of the reference in one HBM pass."""
    lib = _lib.load()
    if not feats.is_cuda:
        raise ValueError("idiaptts_b200 operators need CUDA tensors (there is no CPU fallback)")
    dev = _need_cuda(frame_off, mean, std_dev)
    assert feats.dim() == 2 and feats.dtype == torch.float32 and feats.stride(1) == 1 and frame_off.dtype == torch.int64
    W = feats.shape[1]
    for v in (mean, std_dev):
        assert v is None or (v.dtype == torch.float32 and v.numel() == W)
    if lengths is None:  # one small device-to-host read; pass the host copy of the utterance lengths to stay asynchronous
        lengths = np.diff(frame_off.cpu().numpy())
    U = len(lengths)
    assert U == frame_off.numel() - 1
    t_max = int(lengths.max()) if U else 0
    if min_frames is not None:
        t_max = max(t_max, int(min_frames))
    shape = (U, t_max, W) if batch_first else (t_max, U, W)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    mask = torch.empty(shape[:2] + (1,), dtype=torch.float32, device=dev) if want_mask else None
    with torch.cuda.device(dev):
        check(lib.b2w_pad_normalise(feats.data_ptr(), int(feats.stride(0)) if feats.shape[0] > 1 else W, W, frame_off.data_ptr(), U,
                                    t_max, _ptr(mean), _ptr(std_dev), 1 if batch_first else 0, out.data_ptr(), _ptr(mask),
                                    _stream(dev)), "b2w_pad_normalise")
    return out, mask, lengths


def unpad_denormalise(padded, frame_off, frame_utt, mean=None, std_dev=None, batch_first=False):
    """Inverse of pad_normalise for network outputs: padded [T_max, U, W] (or [U, T_max, W]) -> ragged rows [F, W] * std + mean."""
    lib = _lib.load()
    dev = _need_cuda(padded, frame_off, frame_utt, mean, std_dev)
    assert padded.dim() == 3 and padded.dtype == torch.float32 and frame_off.dtype == torch.int64 and frame_utt.dtype == torch.int32
    U = frame_off.numel() - 1
    t_max = padded.shape[1] if batch_first else padded.shape[0]
    assert (padded.shape[0] if batch_first else padded.shape[1]) == U
    W = padded.shape[2]
    F = frame_utt.numel()
    out = torch.empty((F, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.b2w_unpad_denormalise(padded.data_ptr(), W, frame_off.data_ptr(), frame_utt.data_ptr(), F, U, int(t_max), _ptr(mean),
                                        _ptr(std_dev), 1 if batch_first else 0, out.data_ptr(), W, _stream(dev)),
              "b2w_unpad_denormalise")
    return out


# ----------------------------------------------------------------------------------------------------------------------
# synthesis
# ----------------------------------------------------------------------------------------------------------------------
_randn_tables = {}
_randn_lock = threading.Lock()


def randn_table(n, device):
    """WORLD's randn() stream after randn_reseed(), first n values (cached per device, grown on demand)."""
    lib = _lib.load()
    device = torch.device(device)
    key = str(device)
    with _randn_lock:
        tab = _randn_tables.get(key)
        if tab is None or tab.numel() < n:
            size = max(int(n), 1 << 18)
            tab = torch.empty(size, dtype=torch.float64, device=device)
            with torch.cuda.device(device):
                check(lib.b2w_synth_randn_table(tab.data_ptr(), size, _stream(device)), "b2w_synth_randn_table")
            _randn_tables[key] = tab
        return tab


class SynthPlan:
    """Pulse table of a ragged batch (WORLD GetTimeBase / GetPulseLocationsForTimeBase), launched asynchronously by
    synth_timebase; synth_render turns it into waveforms once the spectral planes exist."""
    pass


def synth_timebase(f0, frame_off, fs, fft_size, frame_period=5.0, status=None, events=None, frame_off_host=None):
    """Launches the pulse-placement kernels (sequential per utterance, a few hundred threads: they overlap well with other
    work on a second stream).  f0 [F] f64, frame_off int64 [U+1] (device).  Returns a SynthPlan.  frame_off_host: the same offsets
    as host numpy; with them the call neither reads from nor (synchronously) writes to the device -- the sample / pulse-slab
    offsets the kernels need are derived from frame_off on the device with the same arithmetic -- so batches can be queued back to
    back (a small device-to-host read would otherwise wait behind the previous batch's waveform copy on the same copy engine)."""
    lib = _lib.load()
    dev = _need_cuda(f0, frame_off)
    assert f0.dtype == torch.float64
    p = SynthPlan()
    foff = frame_off.cpu().numpy() if frame_off_host is None else np.asarray(frame_off_host, np.int64)
    p.U = U = len(foff) - 1
    T = np.diff(foff)
    p.ylen = ylen = (T * frame_period * fs / 1000).astype(np.int64)  # int(T * frame_period * fs / 1000)
    p.out_off = out_off = np.concatenate(([0], np.cumsum(ylen)))
    caps = (ylen.astype(np.float64) * 1200.0 / float(fs)).astype(np.int64) + 64   # = b2w_synth_max_pulses(ylen, fs)
    p.pulse_off = pulse_off = np.concatenate(([0], np.cumsum(caps)))
    if frame_off_host is None:
        p.d_out_off = torch.from_numpy(out_off).to(dev)
        p.d_pulse_off = torch.from_numpy(pulse_off).to(dev)
    else:  # same values, computed on the device (fp64 products of small integers, truncated: identical to the host arithmetic)
        Td = (frame_off[1:] - frame_off[:-1]).double()
        yl = (Td * frame_period * fs / 1000).long()
        zero = torch.zeros(1, dtype=torch.int64, device=dev)
        p.d_out_off = torch.cat((zero, torch.cumsum(yl, 0)))
        p.d_pulse_off = torch.cat((zero, torch.cumsum((yl.double() * 1200.0 / float(fs)).long() + 64, 0)))
    total_cap = int(pulse_off[-1])
    p.pulse_index = torch.empty(max(total_cap, 1), dtype=torch.int32, device=dev)
    p.pulse_shift = torch.empty(max(total_cap, 1), dtype=torch.float64, device=dev)
    p.pulse_vuv = torch.empty(max(total_cap, 1), dtype=torch.uint8, device=dev)
    p.num_pulses = torch.zeros(max(U, 1), dtype=torch.int32, device=dev)
    p.status = new_status(dev) if status is None else status
    p.fs, p.frame_period, p.fft_size, p.frame_off, p.device = int(fs), float(frame_period), int(fft_size), frame_off, dev
    p.empty = U == 0 or out_off[-1] == 0
    if p.empty:
        return p
    p.tab = randn_table(int(ylen.max()) + 1, dev)
    with torch.cuda.device(dev):
        phase_ws = torch.empty(int(out_off[-1]), dtype=torch.float64, device=dev)
        chunk_ws = torch.empty(U * int(lib.b2w_synth_timebase_chunks(int(ylen.max()))), dtype=torch.int32, device=dev)
        _timed(events, "synth_timebase", int(out_off[-1]), lambda: check(lib.b2w_synth_timebase(
            f0.data_ptr(), frame_off.data_ptr(), p.d_out_off.data_ptr(), p.d_pulse_off.data_ptr(), U,
            int(ylen.max()), int(fs), float(frame_period), fft_size, phase_ws.data_ptr(), chunk_ws.data_ptr(),
            p.pulse_index.data_ptr(), p.pulse_shift.data_ptr(), p.pulse_vuv.data_ptr(),
            p.num_pulses.data_ptr(), p.status.data_ptr(), _stream(dev)), "b2w_synth_timebase"))
    return p


def synth_render(p, sp, ap, deemphasis=0.0, out_dtype=torch.float64, debug=None, events=None, precision="f64", sync_counts=True):
    """Minimum-phase responses of every pulse of the plan + overlap-add.  sp, ap [F, K] (f64 or f32, same dtype).
    precision: "f64" = double precision throughout (pyworld-compatible calls); "fast" = one warp per pulse, single-precision
    transforms, float32 responses (the batched Synthesiser path; fft size 1024 only, other sizes fall back to "f64"): same pulse
    decisions, waveform SNR against "f64" ~ 110 dB."""
    assert precision in ("f64", "fast")
    if p.fft_size != 1024:
        precision = "f64"
    lib = _lib.load()
    dev = _need_cuda(sp, ap)
    assert sp.dtype == ap.dtype and sp.dtype in (torch.float32, torch.float64)
    assert 2 * (sp.shape[1] - 1) == p.fft_size
    U, fft_size, out_off, pulse_off = p.U, p.fft_size, p.out_off, p.pulse_off
    y = torch.empty(int(out_off[-1]), dtype=out_dtype, device=dev)
    if p.empty:
        return y, out_off, p.status
    st = _stream(dev)
    with torch.cuda.device(dev):
        fast = precision == "fast"
        if sync_counts or not fast or debug is not None:
            # the response buffer is sized by the ACTUAL pulse counts (one small D2H of U ints)
            npul = p.num_pulses.cpu().numpy()[:U].astype(np.int64)
            max_p = int(npul.max()) if U else 0
            # responses are addressed by the slab offsets, so allocate slab-sized storage only up to the last used row
            last_row = int((pulse_off[:-1] + npul).max()) if U else 0
            total_p = int(npul.sum())
        else:
            # batched path: no read-back.  The slab (an upper bound of the pulse count per utterance) is allocated whole -- the last
            # used row is within one utterance of its end anyway -- and the kernels take the actual counts from device memory.
            npul, max_p, last_row = None, 1, int(pulse_off[-1])
            counts = p.num_pulses
            total_p = lambda: int(counts[:U].sum().item())   # (bench.py's per-kernel table evaluates it after the timed region)
        response = torch.empty((max(last_row, 1), fft_size), dtype=torch.float32 if fast else torch.float64, device=dev)
        if max_p > 0 and fast:
            _timed(events, "render", total_p, lambda: check(lib.b2w_synth_render_f32(
                sp.data_ptr(), ap.data_ptr(), _DT[sp.dtype], p.frame_off.data_ptr(),
                p.d_pulse_off.data_ptr(), p.num_pulses.data_ptr(), U, p.pulse_index.data_ptr(),
                p.pulse_shift.data_ptr(), p.pulse_vuv.data_ptr(), p.tab.data_ptr(), p.tab.numel(), p.fs,
                p.frame_period, fft_size, last_row, response.data_ptr(), st), "b2w_synth_render_f32"))
        elif max_p > 0:
            _timed(events, "render", total_p, lambda: check(lib.b2w_synth_render(
                sp.data_ptr(), ap.data_ptr(), _DT[sp.dtype], p.frame_off.data_ptr(),
                p.d_pulse_off.data_ptr(), p.num_pulses.data_ptr(), U, p.pulse_index.data_ptr(),
                p.pulse_shift.data_ptr(), p.pulse_vuv.data_ptr(), p.tab.data_ptr(), p.tab.numel(), p.fs,
                p.frame_period, fft_size, max_p, response.data_ptr(), st), "b2w_synth_render"))
        if debug is not None:  # diagnostics for the parity tests: the pulse table of every utterance
            debug.update(pulse_off=pulse_off, num_pulses=npul, pulse_index=p.pulse_index, pulse_shift=p.pulse_shift,
                         pulse_vuv=p.pulse_vuv, response=response)
        _timed(events, "overlap_add", total_p, lambda: check((lib.b2w_synth_overlap_add_f32 if fast else lib.b2w_synth_overlap_add)(
            response.data_ptr(), p.d_out_off.data_ptr(), p.d_pulse_off.data_ptr(),
            p.num_pulses.data_ptr(), U, p.pulse_index.data_ptr(), fft_size, int(p.ylen.max()),
            float(deemphasis), y.data_ptr(), _DT[y.dtype], st), "b2w_synth_overlap_add"))
    return y, out_off, p.status


def synthesize(f0, sp, ap, frame_off, fs, frame_period=5.0, deemphasis=0.0, out_dtype=torch.float64, status=None,
               debug=None):
    """pyworld.synthesize on a ragged batch.  f0 [F] f64; sp, ap [F, K] (f64 or f32, same dtype); frame_off int64 [U+1]
    (device) -> (y packed [sum y_len], out_off int64 [U+1] (host numpy))."""
    _need_cuda(f0, sp, ap, frame_off)
    plan = synth_timebase(f0, frame_off, fs, 2 * (sp.shape[1] - 1), frame_period, status)
    return synth_render(plan, sp, ap, deemphasis, out_dtype, debug)


# ----------------------------------------------------------------------------------------------------------------------
# Neural-VTLN all-pass warp
# ----------------------------------------------------------------------------------------------------------------------
def _check_allpass_blocks(blocks):
    """The reference halves / doubles c0 only in the first three n-blocks (static, delta, delta-delta: AllPassWarp.py:162, :171,
    `0:3n:n`); the kernels apply it to every block, which is the same thing up to three blocks -- all the layer is used with."""
    if blocks > 3:
        raise ValueError("all-pass warp of {} n-blocks: the reference's single-sided c0 adaptation covers blocks 0-2 only "
                         "(AllPassWarp.py:162); more than 3 blocks are not supported".format(blocks))


def allpass_forward(x, alpha, n, mean=None, std_dev=None, impl=None):
    """x [rows, blocks*n] f32, alpha [rows] f32 -> y [rows, blocks*n].
    impl: "tc" = tensor-core GEMM for runs of rows that share alpha + recursion for the rest (default when n % 4 == 0, n <= 64),
    "cc" = per-row recursion only."""
    lib = _lib.load()
    dev = _need_cuda(x, alpha, mean, std_dev)
    assert x.dtype == torch.float32 and alpha.dtype == torch.float32 and x.dim() == 2
    rows, width = x.shape
    assert width % n == 0 and alpha.numel() == rows
    _check_allpass_blocks(width // n)
    y = torch.empty_like(x)
    if impl is None:
        impl = "tc" if (n % 4 == 0 and n <= 64 and x.data_ptr() % 16 == 0) else "cc"
    if impl == "tc":
        flags = torch.empty(((rows * (width // n) + 127) // 128,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.b2w_allpass_forward_tc(x.data_ptr(), alpha.data_ptr(), rows, int(n), width // n, _ptr(mean), _ptr(std_dev),
                                             y.data_ptr(), flags.data_ptr(), _stream(dev)), "b2w_allpass_forward_tc")
        return y
    with torch.cuda.device(dev):
        check(lib.b2w_allpass_forward(x.data_ptr(), alpha.data_ptr(), rows, int(n), width // n, _ptr(mean), _ptr(std_dev),
                                      y.data_ptr(), _stream(dev)), "b2w_allpass_forward")
    return y


def allpass_backward(grad_y, x, alpha, n, mean=None, std_dev=None, impl=None):
    """Gradients of allpass_forward w.r.t. x and alpha (impl as in allpass_forward)."""
    lib = _lib.load()
    dev = _need_cuda(grad_y, x, alpha, mean, std_dev)
    rows, width = x.shape
    blocks = width // n
    _check_allpass_blocks(blocks)
    gx = torch.empty_like(x)
    ga = torch.empty_like(alpha)
    ws = torch.empty(rows * blocks, dtype=torch.float32, device=dev)
    if impl is None:
        impl = "tc" if (n % 4 == 0 and n <= 64 and x.data_ptr() % 16 == 0 and grad_y.data_ptr() % 16 == 0) else "cc"
    if impl == "tc":
        flags = torch.empty(((rows * blocks + 127) // 128,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.b2w_allpass_backward_tc(grad_y.data_ptr(), x.data_ptr(), alpha.data_ptr(), rows, int(n), blocks, _ptr(mean),
                                              _ptr(std_dev), gx.data_ptr(), ga.data_ptr(), ws.data_ptr(), flags.data_ptr(),
                                              _stream(dev)), "b2w_allpass_backward_tc")
        return gx, ga
    with torch.cuda.device(dev):
        check(lib.b2w_allpass_backward(grad_y.data_ptr(), x.data_ptr(), alpha.data_ptr(), rows, int(n), blocks, _ptr(mean),
                                       _ptr(std_dev), gx.data_ptr(), ga.data_ptr(), ws.data_ptr(), _stream(dev)),
              "b2w_allpass_backward")
    return gx, ga


# ----------------------------------------------------------------------------------------------------------------------
# MLPG
# ----------------------------------------------------------------------------------------------------------------------
def mlpg(feats, var3, frame_off, D):
    """MLPG.generation on a ragged batch: feats [F, >= 3 D] rows [static | delta | delta-delta] (f32 / f64, last stride 1), var3
    [3 D] f64 (diagonal of the covariance), frame_off int64 [U + 1] -> smoothed trajectories [F, D] f64."""
    lib = _lib.load()
    if not feats.is_cuda:
        raise ValueError("idiaptts_b200 operators need CUDA tensors (there is no CPU fallback)")
    dev = _need_cuda(var3, frame_off)
    assert feats.dim() == 2 and feats.stride(1) == 1 and feats.shape[1] >= 3 * D and feats.dtype in (torch.float32, torch.float64)
    assert var3.dtype == torch.float64 and var3.numel() == 3 * D and frame_off.dtype == torch.int64
    F = feats.shape[0]
    out = torch.empty((F, D), dtype=torch.float64, device=dev)
    ws = torch.empty(int(lib.b2w_mlpg_workspace_doubles(F, int(D))), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.b2w_mlpg(feats.data_ptr(), _DT[feats.dtype], int(feats.stride(0)) if F > 1 else feats.shape[1], var3.data_ptr(),
                           frame_off.data_ptr(), frame_off.numel() - 1, int(D), int(F), ws.data_ptr(), out.data_ptr(), int(D), _stream(dev)),
              "b2w_mlpg")
    return out


# ----------------------------------------------------------------------------------------------------------------------
# objective metrics
# ----------------------------------------------------------------------------------------------------------------------
def world_metrics(org, out, frame_utt, num_utts, num_coded_sps, num_bap):
    """Per-utterance metric sums [num_utts, 8] f64 (see csrc/metrics.cu) of two float32 feature planes [F, D + 2 + nap]."""
    lib = _lib.load()
    dev = _need_cuda(org, out, frame_utt)
    assert org.dtype == torch.float32 and out.dtype == torch.float32 and org.shape == out.shape and frame_utt.dtype == torch.int32
    acc = torch.zeros((num_utts, 8), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.b2w_world_metrics(org.data_ptr(), out.data_ptr(), int(org.shape[1]), frame_utt.data_ptr(), int(org.shape[0]),
                                    int(num_coded_sps), int(num_bap), acc.data_ptr(), _stream(dev)), "b2w_world_metrics")
    return acc
