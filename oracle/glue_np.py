"""CPU oracle (TEST INFRASTRUCTURE, not product code): restatement of the reference's own Python glue
around pyworld/pysptk on the WORLD feature path. Each function cites the reference lines it follows.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import math

import numpy as np

from . import sptk_np, world_np

F0_SILENCE_THRESHOLD = 30  # WorldFeatLabelGen.py:44
LF0_ZERO = 0  # WorldFeatLabelGen.py:45


def preemphasis(raw, p):
    """AudioProcessing.get_raw pre-emphasis, audio/AudioProcessing.py:117-118."""
    raw = np.asarray(raw, np.float64)
    return np.append(raw[0], raw[1:] - p * raw[:-1])


def depreemphasis(raw, p):
    """AudioProcessing.depreemphasis, audio/AudioProcessing.py:330-331: lfilter([1], [1, -p]) i.e. y[n] = x[n] + p*y[n-1]."""
    raw = np.asarray(raw)
    y = np.empty(len(raw), np.float64)
    prev = 0.0
    for i in range(len(raw)):
        prev = float(raw[i]) + p * prev
        y[i] = prev
    return y


def lf0_from_f0(f0, f0_silence_threshold=F0_SILENCE_THRESHOLD, lf0_zero=LF0_ZERO):
    """WorldFeatLabelGen.world_extract_features, world/WorldFeatLabelGen.py:798-799 (float32 log, threshold)."""
    lf0 = np.log(np.asarray(f0, np.float64).clip(min=1e-10), dtype=np.float32)
    lf0[lf0 <= math.log(f0_silence_threshold)] = lf0_zero
    return lf0


def interpolate_lin(data):
    """misc/utils.py:40-86 (Merlin-derived), restated with explicit segment logic but the same arithmetic:
    float32 values, `data[i-1] + step*(k-i+1)` with step computed in float32, the `j < frame_number-1`
    end-of-data rule (a gap whose next voiced frame is the LAST frame is treated as trailing and the last
    frame itself is overwritten too) and the aliasing of ip_data with data."""
    data = np.reshape(np.copy(data), (data.size, 1))
    vuv = np.zeros((data.size, 1))
    vuv[data > 0.0] = 1.0
    n = data.size
    last_value = data.dtype.type(0.0)
    i = 0
    while i < n:
        if data[i, 0] <= 0.0:
            j = i + 1
            found = False
            for j in range(i + 1, n):
                if data[j, 0] > 0.0:
                    found = True
                    break
            if not found and i + 1 >= n:
                j = i + 1  # python leaves j = i + 1 when the range is empty
            if j < n - 1:
                if last_value > 0.0:
                    step = (data[j] - data[i - 1]) / float(j - i)
                    for k in range(i, j):
                        data[k] = data[i - 1] + step * (k - i + 1)
                else:
                    for k in range(i, j):
                        data[k] = data[j]
                # the reference now re-visits i..j-1 as valid data; the net effect is last_value = data[j-1]
                last_value = data[j - 1, 0]
                i = j
            else:
                for k in range(i, n):
                    data[k] = last_value
                # every later frame is now either > 0 (kept, last_value unchanged in effect) or == 0 (refilled with 0)
                break
        else:
            last_value = data[i, 0]
            i += 1
    return data, vuv


def compute_deltas(x):
    """misc/utils.py:103-105: np.gradient along time, cast to float32."""
    return np.gradient(x, axis=0).astype(np.float32)


def world_extract_features(raw, fs, hop_size_ms, f0, n_fft=None, f0_silence_threshold=F0_SILENCE_THRESHOLD,
                           lf0_zero=LF0_ZERO):
    """WorldFeatLabelGen.world_extract_features (world/WorldFeatLabelGen.py:779-807) with the F0 track GIVEN
    (north_star: cached F0; pyworld.wav2world's dio+stonemask stage is outside the path)."""
    raw = np.ascontiguousarray(raw, np.float64)
    T = world_np.num_frames(len(raw), fs, hop_size_ms)
    assert len(f0) == T, (len(f0), T)
    t = world_np.temporal_positions(T, hop_size_ms)
    pow_sp = world_np.cheaptrick(raw, f0, t, fs, fft_size=n_fft)
    ap = world_np.d4c(raw, f0, t, fs, fft_size=n_fft)
    amp_sp = np.sqrt(pow_sp)
    lf0 = lf0_from_f0(f0, f0_silence_threshold, lf0_zero)
    lf0, vuv = interpolate_lin(lf0)
    lf0 = lf0.astype(np.float32)
    vuv = vuv.astype(np.float32)
    bap = np.array(world_np.code_aperiodicity(ap, fs), dtype=np.float32)
    return amp_sp, lf0, vuv, bap


def extract_mcep(amp_sp, num_coded_sps, mgc_alpha):
    """AudioProcessing.extract_mcep, audio/AudioProcessing.py:143-153."""
    mc = sptk_np.mcep(amp_sp, order=num_coded_sps - 1, alpha=mgc_alpha, eps=1.0e-8, min_det=0.0, etype=1, itype=3)
    return mc.astype(np.float32, copy=False)


def mcep_to_amp_sp(mcep, fs, alpha=None, fftlen=None):
    """AudioProcessing.mcep_to_amp_sp, audio/AudioProcessing.py:248-256."""
    if alpha is None:
        alpha = sptk_np.mcepalpha(fs)
    if fftlen is None:
        fftlen = world_np.get_cheaptrick_fft_size(fs)
    sp = sptk_np.mgc2sp(np.ascontiguousarray(mcep, dtype=np.float64), alpha=alpha, gamma=0.0, fftlen=fftlen)
    return np.exp(sp.real.astype(np.float32, copy=False))


def convert_to_world_features(sample, contains_deltas=False, num_coded_sps=60, num_bap=1):
    """WorldFeatLabelGen.convert_to_world_features, world/WorldFeatLabelGen.py:735-762."""
    deltas_factor = 3 if contains_deltas else 1
    num_expected = (num_coded_sps + 1 + num_bap) * deltas_factor + 1
    if sample.shape[1] != num_expected:
        num_expected = (num_coded_sps + 1 + num_bap) * 3 + 1
        if sample.shape[1] == num_expected:
            deltas_factor = 3
            contains_deltas_eff = True
        else:
            raise ValueError("WORLD requires all features to be present.")
    coded_sp = sample[:, :num_coded_sps]
    lf0 = sample[:, num_coded_sps * deltas_factor]
    vuv = np.copy(sample[:, num_coded_sps * deltas_factor + deltas_factor])
    vuv[vuv < 0.5] = 0.0
    vuv[vuv >= 0.5] = 1.0
    if contains_deltas:
        bap = sample[:, -num_bap * 3:-num_bap * 2]
    else:
        bap = sample[:, -num_bap:]
    return coded_sp, lf0, vuv, bap


def world_features_to_raw(amp_sp, lf0, vuv, bap, fs, n_fft=None, f0_silence_threshold=F0_SILENCE_THRESHOLD,
                          lf0_zero=LF0_ZERO, preemphasis_coef=0.0):
    """WorldFeatLabelGen.world_features_to_raw, world/WorldFeatLabelGen.py:910-945."""
    if n_fft is None:
        n_fft = world_np.get_cheaptrick_fft_size(fs)
    pow_sp = np.square(amp_sp, dtype=np.float64)
    f0 = np.exp(lf0, dtype=np.float64)
    vuv = np.array(vuv, copy=True)
    vuv[f0 < f0_silence_threshold] = 0
    f0[vuv == 0] = lf0_zero
    if f0.ndim > 1:
        f0 = f0.squeeze()
    if bap.ndim < 2:
        bap = bap.reshape(-1, 1)
    ap = world_np.decode_aperiodicity(np.ascontiguousarray(bap, np.float64), fs, n_fft)
    raw = world_np.synthesize(f0, pow_sp, ap, fs).astype(np.float32, copy=False)
    return depreemphasis(raw, preemphasis_coef)


class MeanStdDev:
    """misc/normalisation/MeanStdDevExtractor.py:31-53 (add_sample / get_params) incl. its dtype behaviour:
    sums start as python int 0 and become arrays of the sample dtype (float32 for WORLD features)."""

    def __init__(self):
        self.sum_length = 0
        self.sum_frames = 0
        self.sum_squared_frames = 0

    def add_sample(self, sample):
        self.sum_length += len(sample)
        self.sum_frames += np.sum(sample, axis=0)
        self.sum_squared_frames += np.sum(sample ** 2, axis=0)

    def get_params(self):
        mean = self.sum_frames / self.sum_length
        std_dev = np.sqrt(self.sum_squared_frames / self.sum_length - mean ** 2)
        return np.atleast_1d(mean), np.atleast_1d(std_dev)


def combine_mean_std(stats):
    """MeanStdDevExtractor.combine_stats + combine_mean_std, :163-204 and :230-241. stats: iterable of (N, sum, sumsq)."""
    n = 0
    s = 0
    ss = 0
    for cn, cs, css in stats:
        n += cn
        s = s + cs
        ss = ss + css
    mean = s / n
    var = ss / n - mean ** 2
    var = np.where(var < 0, 0.0, var)
    return mean, np.sqrt(var)


def mcd_db(c_ref, c_test):
    """Mel-cepstral distortion as in Metrics.mcd_k (src/Metrics.py:84-92, nnmnkwii.melcd): c0 excluded."""
    d = np.asarray(c_ref, np.float64)[:, 1:] - np.asarray(c_test, np.float64)[:, 1:]
    return float((10.0 / math.log(10.0)) * math.sqrt(2.0) * np.mean(np.sqrt(np.sum(d * d, axis=1))))


# ---------------------------------------------------------------------------------------------------------------
# Neural-VTLN all-pass warp (layers/AllPassWarp.py:148-205); fp64 freqt recursion = oracle also at n = 60 where the
# reference's float32 polynomial tensor overflows.
# ---------------------------------------------------------------------------------------------------------------
def combine_warping_parameters(alphas):
    """AllPassWarp.combine_warping_parameters, layers/AllPassWarp.py:176-184."""
    if isinstance(alphas, (list, tuple)):
        out = alphas[0]
        for a in alphas[1:]:
            out = (out + a) / (1 + out * a)
        return out
    return alphas


def allpass_warp_forward(x, alpha, n):
    """y for x [..., n*blocks] with one alpha per leading index (alpha shape x.shape[:-1] or broadcastable [...,1]).

    Per n-block: x'[0] = x[0]/2; y = x' . W(alpha), W = freqt_matrix(n-1, n-1, alpha)^T; y[0] *= 2."""
    x = np.asarray(x, np.float64)
    lead = x.shape[:-1]
    a = np.broadcast_to(np.asarray(alpha, np.float64).reshape(*np.asarray(alpha).shape[:len(lead)], -1)[..., 0], lead)
    xf = x.reshape(-1, x.shape[-1])
    af = a.reshape(-1)
    out = np.empty_like(xf)
    cache = {}
    for i in range(xf.shape[0]):
        key = float(af[i])
        if key not in cache:
            cache[key] = sptk_np.freqt_matrix(n - 1, n - 1, key)
        A = cache[key]
        for b in range(xf.shape[1] // n):
            v = xf[i, b * n:(b + 1) * n].copy()
            v[0] /= 2.0
            y = A @ v
            y[0] *= 2.0
            out[i, b * n:(b + 1) * n] = y
    return out.reshape(x.shape)


def freqt_with_tangent(x, alpha):
    """freqt(x, len(x)-1, alpha) and its derivative w.r.t. alpha (fp64), by differentiating the recursion."""
    n = len(x)
    a = float(alpha)
    b = 1.0 - a * a
    g = np.zeros(n)
    t = np.zeros(n)
    for r in range(n - 1, -1, -1):
        d, td = g.copy(), t.copy()
        g[0] = x[r] + a * d[0]
        t[0] = d[0] + a * td[0]
        if n > 1:
            g[1] = b * d[0] + a * d[1]
            t[1] = -2.0 * a * d[0] + b * td[0] + d[1] + a * td[1]
        for j in range(2, n):
            g[j] = d[j - 1] + a * (d[j] - g[j - 1])
            t[j] = td[j - 1] + (d[j] - g[j - 1]) + a * (td[j] - t[j - 1])
    return g, t


def allpass_warp_backward(grad_y, x, alpha, n):
    """Gradients of allpass_warp_forward w.r.t. x and alpha (one alpha per row): (grad_x [rows, W], grad_alpha [rows])."""
    x = np.asarray(x, np.float64)
    gy = np.asarray(grad_y, np.float64)
    al = np.asarray(alpha, np.float64).reshape(-1)
    gx = np.zeros_like(x)
    ga = np.zeros(x.shape[0])
    S1 = np.ones(n)
    S1[0] = 0.5
    S2 = np.ones(n)
    S2[0] = 2.0
    for i in range(x.shape[0]):
        A = sptk_np.freqt_matrix(n - 1, n - 1, float(al[i]))
        J = (S2[:, None] * A) * S1[None, :]
        for b in range(x.shape[1] // n):
            sl = slice(b * n, (b + 1) * n)
            gx[i, sl] = J.T @ gy[i, sl]
            _, t = freqt_with_tangent(x[i, sl] * S1, float(al[i]))
            ga[i] += np.dot(gy[i, sl] * S2, t)
    return gx, ga


# ---------------------------------------------------------------------------------------------------------------
# Objective metrics (src/Metrics.py:84-164), restated for the parity test of idiaptts_b200/Metrics.py
# ---------------------------------------------------------------------------------------------------------------
def world_metrics(org_sp, org_lf0, org_vuv, org_bap, out_sp, out_lf0, out_vuv, out_bap):
    T = len(out_sp)
    org_sp, org_lf0, org_vuv, org_bap = (np.asarray(a, np.float64)[:T] for a in (org_sp, org_lf0, org_vuv, org_bap))
    out_sp, out_lf0, out_vuv, out_bap = (np.asarray(a, np.float64) for a in (out_sp, out_lf0, out_vuv, out_bap))
    res = {"MCD": mcd_db(org_sp, out_sp)}                                                                      # :84-92
    res["F0 RMSE"] = math.sqrt((((np.exp(org_lf0) - np.exp(out_lf0)) ** 2) * org_vuv).sum() / org_vuv.sum())   # :94-106
    err20 = np.abs(org_lf0 - out_lf0) > 0.2 * org_lf0
    both = org_vuv * out_vuv
    res["GPE"] = (err20 * both).sum() / both.sum()                                                            # :108-126
    res["VDE"] = (org_vuv != out_vuv).sum() / len(out_vuv)                                                    # :150-155
    res["FFE"] = (err20 * both).sum() / len(out_vuv) + res["VDE"]                                             # :128-148
    if out_bap.ndim > 1 and out_bap.shape[1] > 1:                                                             # :157-164
        res["BAP distortion"] = mcd_db(org_bap, out_bap)
    else:
        res["BAP distortion"] = math.sqrt(((org_bap - out_bap) ** 2).mean()) * (10.0 / math.log(10.0) * math.sqrt(2.0))
    return res


# --------------------------------------------------------------------------------------------------------------
# trainer-facing batch
# --------------------------------------------------------------------------------------------------------------
def prepare_batch(samples, mean=None, std_dev=None, batch_first=False, min_frames=None):
    """WorldFeatLabelGen.preprocess_sample (world/WorldFeatLabelGen.py:279-336: np.float32((sample - mean) / std_dev)) on every
    sample, then ModularModelHandlerPyTorch.prepare_batch (ModularModelHandlerPyTorch.py:389-499): zero-padding to the longest
    sample (torch pad_sequence; min_frames pads further) and sequence_mask (:463-491).  Returns (padded, mask, lengths)."""
    lengths = np.array([len(s) for s in samples], np.int64)
    t_max = int(lengths.max()) if len(samples) else 0
    if min_frames is not None:
        t_max = max(t_max, int(min_frames))
    W = samples[0].shape[1]
    out = np.zeros((len(samples), t_max, W), np.float32)
    mask = np.zeros((len(samples), t_max, 1), np.float32)
    for b, s in enumerate(samples):
        v = s.astype(np.float32)
        if mean is not None:
            v = np.float32((v - np.asarray(mean, np.float32)) / np.asarray(std_dev, np.float32))
        out[b, :len(s)] = v
        mask[b, :len(s)] = 1.0
    if not batch_first:
        out, mask = out.transpose(1, 0, 2).copy(), mask.transpose(1, 0, 2).copy()
    return out, mask, lengths


def unprepare_batch(padded, lengths, mean=None, std_dev=None, batch_first=False):
    """Inverse for network outputs: WorldFeatLabelGen.postprocess_sample's de-normalisation (:338-355: sample * std_dev + mean) of the
    real frames of every batch entry -> list of [T_b, W] float32."""
    if not batch_first:
        padded = padded.transpose(1, 0, 2)
    out = []
    for b, n in enumerate(lengths):
        v = padded[b, :n].astype(np.float32)
        if mean is not None:
            v = np.float32(v * np.asarray(std_dev, np.float32) + np.asarray(mean, np.float32))
        out.append(v)
    return out
