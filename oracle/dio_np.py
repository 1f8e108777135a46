"""CPU oracle (TEST INFRASTRUCTURE, never on the product path) for the F0 stage of pyworld.wav2world:
DIO (pyworld.dio, speed = 1) followed by StoneMask (pyworld.stonemask).  SURVEY.md 8(f) N1.

The reference calls it through pyworld.wav2world (world/WorldFeatLabelGen.py:792) and pyworld.dio / stonemask
(world/LF0LabelGen.py:263-264); pyworld (un-vendored, unpinned, requirements.txt:6) wraps mmorise/World
src/dio.cpp and src/stonemask.cpp.  This file restates their published algorithm in numpy; it is pinned against the
reference's own golden vectors (test/integration/fixtures/WORLD/cmp_mcep20/*.cmp columns 60 / 63, produced by
wav2world on database/wav/*.wav) in tests/test_oracle_dio.py.

Only decimation ratio 1 (pyworld's default speed = 1, the only value the reference uses) is restated."""
import math

import numpy as np

from .world_np import interp1, kLog2, mround, nuttall_window

kCutOff = 50.0
kMaximumValue = 100000.0
kMySafeGuardMinimum = 1e-12
kFloorF0StoneMask = 40.0


def suitable_fft_size(sample):
    """WORLD GetSuitableFFTSize."""
    return int(2 ** (int(math.log(sample) / kLog2) + 1))


def dio_bands(fs, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2.0):
    n = 1 + int(math.log(f0_ceil / f0_floor) / kLog2 * channels_in_octave)
    return [f0_floor * 2.0 ** ((i + 1) / channels_in_octave) for i in range(n)]


def dio_fft_size(x_length, fs, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2.0):
    b0 = dio_bands(fs, f0_floor, f0_ceil, channels_in_octave)[0]
    y_length = 1 + x_length
    return suitable_fft_size(y_length + mround(fs / kCutOff) * 2 + 1 + 4 * int(1.0 + fs / b0 / 2.0))


def low_cut_filter(n, fft_size):
    """WORLD DesignLowCutFilter: delta minus a unit-sum Hann of n taps, centred on sample 0 (circular)."""
    f = np.zeros(fft_size)
    i = np.arange(1, n + 1)
    f[:n] = 0.5 - 0.5 * np.cos(i * 2.0 * math.pi / (n + 1))
    f = -f / f.sum()
    h = (n - 1) // 2
    out = np.zeros(fft_size)
    out[fft_size - h:] = f[:h]
    out[:n - h] = f[h:n]
    out[0] += 1.0
    return out


def spectrum_for_estimation(x, fs, fft_size):
    y_length = len(x) + 1
    y = np.zeros(fft_size)
    y[:len(x)] = x
    y[:y_length] -= y[:y_length].sum() / y_length
    spec = np.fft.rfft(y)
    cutoff = mround(fs / kCutOff)
    return spec * np.fft.rfft(low_cut_filter(cutoff * 2 + 1, fft_size))


def filtered_signal(half_average_length, fft_size, y_spectrum, y_length):
    lpf = np.zeros(fft_size)
    lpf[:half_average_length * 4] = nuttall_window(half_average_length * 4)
    sig = np.fft.irfft(y_spectrum * np.fft.rfft(lpf), fft_size) * fft_size  # FFTW backward transforms are unnormalised
    bias = half_average_length * 2
    return sig[bias:bias + y_length].copy()


def zero_crossing_engine(sig, fs):
    """negative-going zero crossings -> (interval_locations, intervals)."""
    idx = np.nonzero((sig[:-1] > 0.0) & (sig[1:] <= 0.0))[0] + 1
    if len(idx) < 2:
        return np.zeros(0), np.zeros(0)
    fine = idx - sig[idx - 1] / (sig[idx] - sig[idx - 1])
    intervals = fs / (fine[1:] - fine[:-1])
    locations = (fine[:-1] + fine[1:]) / 2.0 / fs
    return locations, intervals


def four_zero_crossing_intervals(sig, fs):
    out = [zero_crossing_engine(sig, fs), zero_crossing_engine(-sig, fs)]
    d = sig[:-1] - sig[1:]
    out += [zero_crossing_engine(d, fs), zero_crossing_engine(-d, fs)]
    return out


def f0_candidate_contour(events, boundary_f0, f0_floor, f0_ceil, t):
    T = len(t)
    if any(len(loc) - 2 <= 0 for loc, _ in events):
        return np.zeros(T), np.full(T, kMaximumValue)
    s = np.stack([interp1(loc, itv, t) for loc, itv in events])
    cand = (s[0] + s[1] + s[2] + s[3]) / 4.0
    score = np.sqrt(((s[0] - cand) ** 2 + (s[1] - cand) ** 2 + (s[2] - cand) ** 2 + (s[3] - cand) ** 2) / 3.0)
    bad = (cand > boundary_f0) | (cand < boundary_f0 / 2.0) | (cand > f0_ceil) | (cand < f0_floor)
    cand[bad] = 0.0
    score[bad] = kMaximumValue
    return cand, score


def dio_candidates(x, fs, t, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2.0):
    bands = dio_bands(fs, f0_floor, f0_ceil, channels_in_octave)
    y_length = len(x) + 1
    fft_size = dio_fft_size(len(x), fs, f0_floor, f0_ceil, channels_in_octave)
    spec = spectrum_for_estimation(x, fs, fft_size)
    cands, scores = [], []
    for b in bands:
        sig = filtered_signal(mround(fs / b / 2.0), fft_size, spec, y_length)
        c, s = f0_candidate_contour(four_zero_crossing_intervals(sig, fs), b, f0_floor, f0_ceil, t)
        cands.append(c)
        scores.append(s / (c + kMySafeGuardMinimum))
    return np.stack(cands), np.stack(scores)


def best_f0_contour(cands, scores):
    best = cands[0].copy()
    tmp = scores[0].copy()
    for j in range(1, len(cands)):
        m = tmp > scores[j]
        tmp[m] = scores[j][m]
        best[m] = cands[j][m]
    return best


def _select_best_f0(current_f0, past_f0, cands, target, allowed_range):
    ref = (current_f0 * 3.0 - past_f0) / 2.0
    best = cands[0][target]
    err = abs(ref - best)
    for i in range(1, len(cands)):
        e = abs(ref - cands[i][target])
        if e < err:
            err, best = e, cands[i][target]
    if abs(1.0 - best / ref) > allowed_range:
        return 0.0
    return best


def _boundary_list(f0):
    vuv = (f0 > 0).astype(np.int64)
    vuv[0] = vuv[-1] = 0
    out = []
    for i in range(len(f0) - 1):
        if vuv[i + 1] - vuv[i] != 0:
            out.append(i + len(out) % 2)
    return out


def fix_f0_contour(frame_period, cands, best, f0_floor, allowed_range, step2="erosion"):
    """WORLD FixF0Contour.  step2 = "erosion" is the variant the reference's fixtures were produced with (a frame survives
    only if the voice_range_minimum frames centred on it are voiced; it reproduces all 9 golden utterances, vuv bit-exact);
    step2 = "sections" is the later WORLD variant that only drops voiced sections shorter than voice_range_minimum (it
    disagrees with the fixtures on 3-69 frames per utterance)."""
    T = len(best)
    vrm = int(0.5 + 1000.0 / frame_period / f0_floor) * 2 + 1
    if T <= vrm:
        return np.zeros(T)
    # step 1: reject jumps
    base = np.zeros(T)
    base[vrm:T - vrm] = best[vrm:T - vrm]
    s1 = np.zeros(T)
    for i in range(vrm, T):
        s1[i] = base[i] if abs((base[i] - base[i - 1]) / (kMySafeGuardMinimum + base[i])) < allowed_range else 0.0
    # step 2: erosion -- a frame survives only if the voice_range_minimum frames centred on it are all voiced
    s2 = s1.copy()
    if step2 == "erosion":
        center = (vrm - 1) // 2
        for i in range(center, T - center):
            if np.any(s1[i - center:i + center + 1] == 0):
                s2[i] = 0.0
    else:
        bl = _boundary_list(s1)
        for i in range(len(bl) // 2):
            if bl[2 * i + 1] - bl[2 * i] < vrm:
                s2[bl[2 * i]:bl[2 * i + 1] + 1] = 0.0
    pos, neg = [], []
    for i in range(1, T):
        if s2[i] == 0 and s2[i - 1] != 0:
            neg.append(i - 1)
        elif s2[i - 1] == 0 and s2[i] != 0:
            pos.append(i)
    # step 3: extend forward
    s3 = s2.copy()
    for i, n0 in enumerate(neg):
        limit = T - 1 if i == len(neg) - 1 else neg[i + 1]
        for j in range(n0, limit):
            s3[j + 1] = _select_best_f0(s3[j], s3[j - 1], cands, j + 1, allowed_range)
            if s3[j + 1] == 0:
                break
    # step 4: extend backward
    s4 = s3.copy()
    for i in range(len(pos) - 1, -1, -1):
        limit = 1 if i == 0 else pos[i - 1]
        for j in range(pos[i], limit, -1):
            s4[j - 1] = _select_best_f0(s4[j], s4[j + 1], cands, j - 1, allowed_range)
            if s4[j - 1] == 0:
                break
    return s4


def dio(x, fs, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2.0, frame_period=5.0, speed=1, allowed_range=0.1):
    """pyworld.dio -> (f0 [T], t [T])."""
    if speed != 1:
        raise NotImplementedError("only speed = 1 (pyworld default) is restated")
    x = np.asarray(x, dtype=np.float64)
    T = int(1000.0 * len(x) / fs / frame_period) + 1
    t = np.arange(T) * frame_period / 1000.0
    cands, scores = dio_candidates(x, fs, t, f0_floor, f0_ceil, channels_in_octave)
    best = best_f0_contour(cands, scores)
    return fix_f0_contour(frame_period, cands, best, f0_floor, allowed_range), t


# --------------------------------------------------------------------------------------------------------------
# StoneMask
# --------------------------------------------------------------------------------------------------------------
def _fix_f0(power, numer, fft_size, fs, initial_f0, nharm):
    num = den = 0.0
    for i in range(nharm):
        idx = mround(initial_f0 * fft_size / fs * (i + 1))
        k = idx % fft_size  # WORLD reads past fft_size / 2 when 6 * f0 > fs / 2 (undefined there); the spectrum is periodic
        inst = 0.0 if power[k] == 0.0 else idx * fs / fft_size + numer[k] / power[k] * fs / 2.0 / math.pi
        amp = math.sqrt(power[k])
        num += amp * inst
        den += amp * (i + 1.0)
    return num / (den + kMySafeGuardMinimum)


def stonemask_frame(x, fs, pos, initial_f0):
    if initial_f0 <= kFloorF0StoneMask or initial_f0 > fs / 12.0:
        return 0.0
    half = int(1.5 * fs / initial_f0 + 1.0)
    wlen_t = (2.0 * half + 1.0) / fs
    base_time = (np.arange(2 * half + 1) - half) / fs
    fft_size = int(2 ** (2.0 + int(math.log(half * 2.0 + 1.0) / kLog2)))
    index_raw = np.array([mround(v) for v in (pos + base_time) * fs], dtype=np.int64)
    tmp = (index_raw - 1.0) / fs - pos
    w = 0.42 + 0.5 * np.cos(2.0 * math.pi * tmp / wlen_t) + 0.08 * np.cos(4.0 * math.pi * tmp / wlen_t)
    dw = np.empty_like(w)
    dw[0] = -w[1] / 2.0
    dw[1:-1] = -(w[2:] - w[:-2]) / 2.0
    dw[-1] = w[-2] / 2.0
    seg = x[np.clip(index_raw - 1, 0, len(x) - 1)]
    main = np.fft.fft(seg * w, fft_size)
    diff = np.fft.fft(seg * dw, fft_size)
    power = main.real ** 2 + main.imag ** 2
    numer = main.real * diff.imag - main.imag * diff.real
    tentative = _fix_f0(power, numer, fft_size, fs, initial_f0, 2)
    if tentative <= 0.0 or tentative > initial_f0 * 2:
        mean_f0 = 0.0
    else:
        mean_f0 = _fix_f0(power, numer, fft_size, fs, tentative, 6)
    if abs(mean_f0 - initial_f0) > initial_f0 * 0.2:
        mean_f0 = initial_f0
    return mean_f0


def stonemask(x, f0, t, fs):
    """pyworld.stonemask -> refined f0 [T]."""
    x = np.asarray(x, dtype=np.float64)
    return np.array([stonemask_frame(x, fs, t[i], f0[i]) for i in range(len(f0))])


def wav2world_f0(x, fs, frame_period=5.0):
    """The F0 half of pyworld.wav2world (WorldFeatLabelGen.py:792)."""
    f0, t = dio(x, fs, frame_period=frame_period)
    return stonemask(x, f0, t, fs), t
