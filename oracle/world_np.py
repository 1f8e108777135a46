"""CPU oracle (TEST INFRASTRUCTURE, not product code): numpy restatement of the WORLD vocoder stages
that IdiapTTS reaches through pyworld.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (idiaptts_b200/) never does.

pyworld (PyPI, un-pinned, reference requirements.txt:6) wraps mmorise/World; neither is present under
/root/reference, so the published algorithm is restated here in fp64 and anchored on the reference's
own call sites and golden fixtures:

  * call sites   idiaptts/src/data_preparation/world/WorldFeatLabelGen.py:792 (wav2world -> cheaptrick, d4c),
                 :805 (code_aperiodicity), :940 (decode_aperiodicity), :943 (synthesize)
                 idiaptts/src/data_preparation/audio/AudioProcessing.py:60,:71 (fft size, #bands)
  * goldens      test/integration/fixtures/WORLD/cmp_mcep20/*.cmp  (analysis half: PINNED, see
                 tests/test_oracle_golden.py - bap to 2e-5, mcep via oracle/sptk_np.py to 2e-6)
  * synthesis    PARITY UNPINNED: the reference holds no golden waveform; `synthesize`
                 below is a careful restatement of WORLD's Synthesis() and defines parity for our CUDA path.

WORLD's two "safe-guard" noise injections (randn()*1e-12 added to every windowed sample and
eps*|randn()| added to the smoothed spectrum) are replaced by 0 and +eps: they are below every stated
tolerance and un-reproducible without WORLD's global RNG call order.
"""
import math

import numpy as np

kPi = 3.1415926535897932384
kMySafeGuardMinimum = 1e-12
kEps = 2.2204460492503131e-16
kDefaultF0 = 500.0
kFloorF0 = 71.0
kLog2 = 0.69314718055994529
kFrequencyInterval = 3000.0
kUpperLimit = 15000.0
kFloorF0D4C = 47.0
kThresholdD4C = 0.85
kQ1 = -0.15
default_frame_period = 5.0


def mround(x):
    """MATLAB round (half away from zero), WORLD matlab_round."""
    return int(x + 0.5) if x > 0 else int(x - 0.5)


def get_cheaptrick_fft_size(fs, f0_floor=kFloorF0):
    """WORLD GetFFTSizeForCheapTrick; reference use: AudioProcessing.py:53-60."""
    return int(2 ** (1.0 + int(math.log(3.0 * fs / f0_floor + 1) / kLog2)))


def get_num_aperiodicities(fs):
    """WORLD GetNumberOfAperiodicities; reference use: AudioProcessing.py:70-71."""
    return int(min(kUpperLimit, fs / 2.0 - kFrequencyInterval) / kFrequencyInterval)


def get_d4c_fft_size(fs):
    return int(2 ** (1.0 + int(math.log(4.0 * fs / kFloorF0D4C + 1) / kLog2)))


def get_lovetrain_fft_size(fs):
    return int(2 ** (1.0 + int(math.log(3.0 * fs / 40.0 + 1) / kLog2)))


def num_frames(x_length, fs, frame_period=default_frame_period):
    """WORLD GetSamplesForDIO: the frame count wav2world produces (WorldFeatLabelGen.py:792)."""
    return int(1000.0 * x_length / fs / frame_period) + 1


def temporal_positions(n, frame_period=default_frame_period):
    return np.arange(n) * frame_period / 1000.0


# --------------------------------------------------------------------------------------------------------------
# shared helpers (WORLD common.cpp / matlabfunctions.cpp)
# --------------------------------------------------------------------------------------------------------------
def interp1Q(x0, dx, y, xi):
    """Linear interpolation on a uniform grid starting at x0 with step dx (dx may be negative)."""
    pos = (xi - x0) / dx
    base = pos.astype(np.int64)  # truncation toward zero, like static_cast<int>
    frac = pos - base
    dy = np.empty_like(y)
    dy[:-1] = y[1:] - y[:-1]
    dy[-1] = 0.0
    return y[base] + dy[base] * frac


def interp1(x, y, xi):
    """WORLD interp1 (histc based): k with x[k-1] <= xi < x[k]; yi = y[k-1] + s*(y[k]-y[k-1])."""
    k = np.searchsorted(x, xi, side="right")
    k = np.clip(k, 1, len(x) - 1)
    s = (xi - x[k - 1]) / (x[k] - x[k - 1])
    return y[k - 1] + s * (y[k] - y[k - 1])


def dc_correction(p, f0, fs, fft_size):
    upper = 2 + int(f0 * fft_size / fs)
    axis = np.arange(upper) * fs / fft_size
    replica = interp1Q(f0 - axis[0], -float(fs) / fft_size, p[:upper + 1], axis[:upper - 1])
    out = p.copy()
    out[:upper - 1] += replica
    return out


def linear_smoothing(p, width, fs, fft_size):
    b = int(width * fft_size / fs) + 1
    h = fft_size // 2
    mir = np.concatenate((p[b:0:-1], p[:h], p[h:h - b - 1:-1]))  # length h + 2b + 1
    seg = np.cumsum(mir * fs / fft_size)
    fax = np.arange(h + 1) / fft_size * fs - width / 2.0
    origin = -(b - 0.5) * fs / fft_size
    dfi = float(fs) / fft_size
    low = interp1Q(origin, dfi, seg, fax)
    high = interp1Q(origin, dfi, seg, fax + width)
    return (high - low) / width


def nuttall_window(n):
    tmp = np.arange(n) / (n - 1.0)
    return (0.355768 - 0.487396 * np.cos(2.0 * kPi * tmp) + 0.144232 * np.cos(4.0 * kPi * tmp)
            - 0.012604 * np.cos(6.0 * kPi * tmp))


# --------------------------------------------------------------------------------------------------------------
# CheapTrick (WORLD cheaptrick.cpp)
# --------------------------------------------------------------------------------------------------------------
def _cheaptrick_frame(x, fs, f0, pos, fft_size, q1):
    half = mround(1.5 * fs / f0)
    base = np.arange(-half, half + 1)
    origin = mround(pos * fs + 0.001)
    idx = np.clip(origin + base, 0, len(x) - 1)
    position = base / 1.5 / fs
    win = 0.5 * np.cos(kPi * position * f0) + 0.5
    win = win / math.sqrt(np.sum(win * win))
    w = x[idx] * win
    w = w - win * (np.sum(w) / np.sum(win))
    spec = np.fft.rfft(w, fft_size)
    p = spec.real ** 2 + spec.imag ** 2
    p = dc_correction(p, f0, fs, fft_size)
    p = linear_smoothing(p, f0 * 2.0 / 3.0, fs, fft_size)
    p = p + kEps
    # smoothing with recovery (cepstral liftering)
    h = fft_size // 2
    quef = np.arange(1, h + 1) / float(fs)
    smooth = np.ones(h + 1)
    comp = np.ones(h + 1)
    smooth[1:] = np.sin(kPi * f0 * quef) / (kPi * f0 * quef)
    comp[0] = (1.0 - 2.0 * q1) + 2.0 * q1
    comp[1:] = (1.0 - 2.0 * q1) + 2.0 * q1 * np.cos(2.0 * kPi * quef * f0)
    cep = np.fft.rfft(np.concatenate((np.log(p), np.log(p[h - 1:0:-1])))).real  # FFT of the mirrored log spectrum
    lift = cep * smooth * comp / fft_size
    env = np.fft.hfft(lift, fft_size)[:h + 1]  # unnormalised c2r of a real (zero-phase) spectrum
    return np.exp(env)


def cheaptrick(x, f0, t, fs, q1=kQ1, f0_floor=kFloorF0, fft_size=None):
    """pyworld.cheaptrick(x, f0, temporal_positions, fs, q1=-0.15, f0_floor=71.0, fft_size=None) -> sp [T, K] (power)."""
    x = np.ascontiguousarray(x, np.float64)
    if fft_size is None:
        fft_size = get_cheaptrick_fft_size(fs, f0_floor)
    floor = 3.0 * fs / (fft_size - 3.0)
    out = np.empty((len(f0), fft_size // 2 + 1))
    for i in range(len(f0)):
        cur = kDefaultF0 if f0[i] <= floor else f0[i]
        out[i] = _cheaptrick_frame(x, fs, cur, t[i], fft_size, q1)
    return out


# --------------------------------------------------------------------------------------------------------------
# D4C (WORLD d4c.cpp), including the LoveTrain voiced/unvoiced stage
# --------------------------------------------------------------------------------------------------------------
def _d4c_window(x, fs, f0, pos, window_type, ratio):
    half = mround(ratio * fs / f0 / 2.0)
    base = np.arange(-half, half + 1)
    origin = mround(pos * fs + 0.001)
    idx = np.clip(origin + base, 0, len(x) - 1)
    position = (2.0 * base / ratio) / fs
    if window_type == "hann":
        win = 0.5 * np.cos(kPi * position * f0) + 0.5
    else:
        win = 0.42 + 0.5 * np.cos(kPi * position * f0) + 0.08 * np.cos(kPi * position * f0 * 2)
    w = x[idx] * win
    return w - win * (np.sum(w) / np.sum(win))


def d4c_lovetrain(x, fs, f0, t):
    """aperiodicity0 per frame (0 for f0 == 0)."""
    nl = get_lovetrain_fft_size(fs)
    b0 = int(math.ceil(100.0 * nl / fs))
    b1 = int(math.ceil(4000.0 * nl / fs))
    b2 = int(math.ceil(7900.0 * nl / fs))
    out = np.zeros(len(f0))
    for i in range(len(f0)):
        if f0[i] == 0.0:
            continue
        cur = max(f0[i], 40.0)
        w = _d4c_window(x, fs, cur, t[i], "blackman", 3.0)
        s = np.fft.rfft(w, nl)
        p = s.real ** 2 + s.imag ** 2
        p[:b0 + 1] = 0.0
        c = np.cumsum(p[:b2 + 1])
        out[i] = c[b1] / c[b2]
    return out


def _d4c_centroid(x, fs, f0, pos, n4):
    w = np.zeros(n4)
    seg = _d4c_window(x, fs, f0, pos, "blackman", 4.0)
    w[:len(seg)] = seg
    n = mround(2.0 * fs / f0) * 2 + 1
    w[:n] = w[:n] / math.sqrt(np.sum(w[:n] * w[:n]))
    s1 = np.fft.rfft(w)
    s2 = np.fft.rfft(w * np.arange(1.0, n4 + 1.0))
    return s2.real * s1.real + s1.imag * s2.imag


def _d4c_frame_coarse(x, fs, f0, pos, n4, nap, window):
    """Coarse aperiodicity in dB (nap values) of one voiced frame; f0 already floored at kFloorF0D4C."""
    wl = len(window)
    sc = _d4c_centroid(x, fs, f0, pos - 0.25 / f0, n4) + _d4c_centroid(x, fs, f0, pos + 0.25 / f0, n4)
    sc = dc_correction(sc, f0, fs, n4)
    s = np.fft.rfft(_d4c_window(x, fs, f0, pos, "hann", 4.0), n4)
    sps = s.real ** 2 + s.imag ** 2
    sps = dc_correction(sps, f0, fs, n4)
    sps = linear_smoothing(sps, f0, fs, n4)
    gd = sc / sps
    gd = linear_smoothing(gd, f0 / 2.0, fs, n4)
    gd = gd - linear_smoothing(gd, f0, fs, n4)
    boundary = mround(n4 * 8.0 / wl)
    half = wl // 2
    coarse = np.empty(nap)
    for i in range(nap):
        center = int(kFrequencyInterval * (i + 1) * n4 / fs)
        s = np.fft.rfft(gd[center - half:center + half + 1] * window, n4)
        p = np.sort(s.real ** 2 + s.imag ** 2)
        c = np.cumsum(p)
        coarse[i] = 10 * math.log10(c[n4 // 2 - boundary - 1] / c[n4 // 2])
    return np.minimum(0.0, coarse + (f0 - 100) / 50.0)


def _coarse_to_aperiodicity(coarse_db, fs, fft_size):
    nap = len(coarse_db)
    cax = np.concatenate((np.arange(nap + 1) * kFrequencyInterval, [fs / 2.0]))
    cap = np.concatenate(([-60.0], coarse_db, [-kMySafeGuardMinimum]))
    fax = np.arange(fft_size // 2 + 1) * float(fs) / fft_size
    return np.power(10.0, interp1(cax, cap, fax) / 20.0)


def d4c_coarse(x, f0, t, fs, threshold=kThresholdD4C):
    """Returns (voiced_mask[T] bool, coarse_db[T, nap]); rows of unvoiced frames are undefined (0)."""
    x = np.ascontiguousarray(x, np.float64)
    n4 = get_d4c_fft_size(fs)
    nap = get_num_aperiodicities(fs)
    wl = int(kFrequencyInterval * n4 / fs) * 2 + 1
    window = nuttall_window(wl)
    ap0 = d4c_lovetrain(x, fs, f0, t)
    voiced = np.zeros(len(f0), bool)
    coarse = np.zeros((len(f0), nap))
    for i in range(len(f0)):
        if f0[i] == 0 or ap0[i] <= threshold:
            continue
        voiced[i] = True
        coarse[i] = _d4c_frame_coarse(x, fs, max(kFloorF0D4C, f0[i]), t[i], n4, nap, window)
    return voiced, coarse


def d4c(x, f0, t, fs, threshold=kThresholdD4C, fft_size=None):
    """pyworld.d4c(x, f0, temporal_positions, fs, threshold=0.85, fft_size=None) -> ap [T, K]."""
    if fft_size is None:
        fft_size = get_cheaptrick_fft_size(fs)
    voiced, coarse = d4c_coarse(x, f0, t, fs, threshold)
    out = np.full((len(f0), fft_size // 2 + 1), 1.0 - kMySafeGuardMinimum)
    for i in np.nonzero(voiced)[0]:
        out[i] = _coarse_to_aperiodicity(coarse[i], fs, fft_size)
    return out


# --------------------------------------------------------------------------------------------------------------
# codec (WORLD codec.cpp)
# --------------------------------------------------------------------------------------------------------------
def code_aperiodicity(ap, fs):
    """pyworld.code_aperiodicity(aperiodicity, fs) -> bap [T, nap] (dB)."""
    ap = np.ascontiguousarray(ap, np.float64)
    fft_size = (ap.shape[1] - 1) * 2
    nap = get_num_aperiodicities(fs)
    cax = kFrequencyInterval * (np.arange(nap) + 1.0)
    out = np.empty((ap.shape[0], nap))
    for i in range(ap.shape[0]):
        out[i] = interp1Q(0.0, float(fs) / fft_size, 20 * np.log10(ap[i]), cax)
    return out


def decode_aperiodicity(coded_ap, fs, fft_size):
    """pyworld.decode_aperiodicity(coded_aperiodicity, fs, fft_size) -> ap [T, K]."""
    coded_ap = np.ascontiguousarray(coded_ap, np.float64)
    out = np.full((coded_ap.shape[0], fft_size // 2 + 1), 1.0 - kMySafeGuardMinimum)
    nap = get_num_aperiodicities(fs)
    for i in range(coded_ap.shape[0]):
        tmp = 0.0
        for j in range(nap):
            tmp += coded_ap[i, j]
        tmp /= nap
        if tmp > -0.5:
            continue
        out[i] = _coarse_to_aperiodicity(coded_ap[i, :nap], fs, fft_size)
    return out


# --------------------------------------------------------------------------------------------------------------
# synthesis (WORLD synthesis.cpp); PARITY UNPINNED, see module docstring
# --------------------------------------------------------------------------------------------------------------
def xorshift_randn_sequence(n):
    """First n values of WORLD's randn() after randn_reseed(): xorshift128, sum of 12 draws (>>4), /2^28 - 6."""
    x, y, z, w = 123456789, 362436069, 521288629, 88675123
    out = np.empty(n)
    M = 0xFFFFFFFF
    for i in range(n):
        tmp = 0
        for _ in range(12):
            t = (x ^ (x << 11)) & M
            x, y, z = y, z, w
            w = ((w ^ (w >> 19)) ^ (t ^ (t >> 8))) & M
            tmp += w >> 4
        out[i] = tmp / 268435456.0 - 6.0
    return out


def get_dc_remover(fft_size):
    h = fft_size // 2
    r = 0.5 - 0.5 * np.cos(2.0 * kPi * (np.arange(h) + 1.0) / (1.0 + fft_size))
    dc = 0.0
    for v in r:  # sequential sum like the reference loop
        dc += v * 2.0
    r = r / dc
    return np.concatenate((r, r[::-1]))


def minimum_phase_spectrum(log_spec_half, fft_size):
    """WORLD GetMinimumPhaseSpectrum: log-amplitude (K bins) -> complex minimum-phase spectrum (K bins)."""
    h = fft_size // 2
    full = np.concatenate((log_spec_half, log_spec_half[h - 1:0:-1]))
    cep = np.fft.fft(full)
    cep = np.conj(cep)
    fold = np.zeros(fft_size, complex)
    fold[0] = cep[0]
    fold[1:h] = 2.0 * cep[1:h]
    fold[h] = cep[h]
    sp = np.fft.fft(fold)[:h + 1] / fft_size
    return np.exp(sp.real) * (np.cos(sp.imag) + 1j * np.sin(sp.imag))


def synthesis_time_base(f0, fs, frame_period_s, y_length, fft_size):
    """Pulse placement: returns (pulse sample index[P], fractional time shift[P] in s, per-sample vuv[y_length])."""
    T = len(f0)
    lowest_f0 = fs // fft_size + 1.0
    coarse_t = np.arange(T + 1) * frame_period_s
    cf0 = np.where(f0 < lowest_f0, 0.0, f0).astype(np.float64)
    cvuv = np.where(cf0 == 0.0, 0.0, 1.0)
    cf0 = np.append(cf0, cf0[-1] * 2 - cf0[-2])
    cvuv = np.append(cvuv, cvuv[-1] * 2 - cvuv[-2])
    time_axis = np.arange(y_length) / float(fs)
    if0 = interp1(coarse_t, cf0, time_axis)
    ivuv = interp1(coarse_t, cvuv, time_axis)
    ivuv = np.where(ivuv > 0.5, 1.0, 0.0)
    if0 = np.where(ivuv == 0.0, kDefaultF0, if0)
    two_pi = 2.0 * kPi
    total = np.cumsum(two_pi * if0 / fs)  # np.cumsum is a sequential fp64 accumulation, like the reference loop
    wrap = np.fmod(total, two_pi)
    jumps = np.abs(wrap[1:] - wrap[:-1])
    idx = np.nonzero(jumps > kPi)[0]
    y1 = wrap[idx] - two_pi
    y2 = wrap[idx + 1]
    shift = (-y1 / (y2 - y1)) / fs
    return idx, shift, ivuv


def synthesize(f0, sp, ap, fs, frame_period=default_frame_period, responses=None):
    """pyworld.synthesize(f0, spectrogram, aperiodicity, fs, frame_period=5.0) -> y [int(T*frame_period*fs/1000)]."""
    f0 = np.ascontiguousarray(f0, np.float64)
    sp = np.ascontiguousarray(sp, np.float64)
    ap = np.ascontiguousarray(ap, np.float64)
    T = len(f0)
    fft_size = (sp.shape[1] - 1) * 2
    h = fft_size // 2
    y_length = int(T * frame_period * fs / 1000)
    fp = frame_period / 1000.0
    y = np.zeros(y_length)
    idx, shift, ivuv = synthesis_time_base(f0, fs, fp, y_length, fft_size)
    P = len(idx)
    if P == 0:
        return y
    dc_remover = get_dc_remover(fft_size)
    total_noise = int(idx[-1] - idx[0])
    randn_seq = xorshift_randn_sequence(total_noise)
    k = np.arange(h + 1)
    for p in range(P):
        n_p = int(idx[p])
        noise_size = int(idx[min(P - 1, p + 1)] - n_p)
        cur_t = n_p / float(fs)
        fl = min(T - 1, int(math.floor(cur_t / fp)))
        ce = min(T - 1, int(math.ceil(cur_t / fp)))
        w = cur_t / fp - fl
        sa_fl = np.clip(ap[fl], 0.001, 0.999999999999) ** 2
        if fl == ce:
            env = np.abs(sp[fl])
            ar = sa_fl
        else:
            env = (1.0 - w) * np.abs(sp[fl]) + w * np.abs(sp[ce])
            ar = (1.0 - w) * sa_fl + w * np.clip(ap[ce], 0.001, 0.999999999999) ** 2
        cur_vuv = ivuv[n_p]
        # periodic response
        if cur_vuv <= 0.5 or ar[0] > 0.999:
            periodic = np.zeros(fft_size)
        else:
            X = minimum_phase_spectrum(np.log(env * (1.0 - ar) + kMySafeGuardMinimum) / 2.0, fft_size)
            coef = 2.0 * kPi * shift[p] * fs / fft_size
            re2 = np.cos(coef * k)
            im2 = np.sqrt(1.0 - re2 * re2)
            X = (X.real * re2 + X.imag * im2) + 1j * (X.imag * re2 - X.real * im2)
            wave = np.fft.irfft(X, fft_size) * fft_size  # unnormalised c2r
            periodic = np.concatenate((wave[h:], wave[:h]))  # fftshift
            dc = np.sum(periodic[h:])
            periodic[:h] = -dc * dc_remover[:h]
            periodic[h:] -= dc * dc_remover[h:]
        # aperiodic response
        start = n_p - int(idx[0])
        noise = np.zeros(fft_size)
        if noise_size > 0:
            seg = randn_seq[start:start + noise_size]
            seg = seg - np.sum(seg) / noise_size
            # noise_size > fft_size (f0 interpolated below fs/fft_size at a voicing edge) overruns WORLD's
            # buffer (undefined behaviour there); we keep the first fft_size samples.
            noise[:min(noise_size, fft_size)] = seg[:fft_size]
        Z = np.fft.rfft(noise)
        if cur_vuv != 0.0:
            X = minimum_phase_spectrum(np.log(env * ar) / 2.0, fft_size)
        else:
            X = minimum_phase_spectrum(np.log(env) / 2.0, fft_size)
        wave = np.fft.irfft(X * Z, fft_size) * fft_size
        aperiodic = np.concatenate((wave[h:], wave[:h]))
        response = (periodic * math.sqrt(float(noise_size)) + aperiodic) / fft_size
        if responses is not None:  # diagnostics for the parity tests
            responses.append((periodic.copy(), aperiodic.copy(), response.copy()))
        offset = n_p - h + 1
        lo = max(0, -offset)
        hi = min(fft_size, y_length - offset)
        y[lo + offset:hi + offset] += response[lo:hi]
    return y
