"""ctypes access to the C oracle (oracle/c/world_oracle.c).  TEST INFRASTRUCTURE / CPU baseline only: may be imported by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by the product path."""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = ctypes.CDLL(_build.build())
        P, I, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        l.oracle_cheaptrick.argtypes = [P, I, I, P, P, I, I, D, P]
        l.oracle_d4c_coarse.argtypes = [P, I, I, P, P, I, D, P, P]
        l.oracle_bap_from_coarse.argtypes = [P, P, I, I, I, P]
        l.oracle_bap_from_coarse.restype = None
        l.oracle_mcep.argtypes = [P, I, I, I, I, D, I, I, D, D, P, P]
        l.oracle_lf0_vuv.argtypes = [P, I, D, ctypes.c_float, P, P]
        l.oracle_lf0_vuv.restype = None
        l.oracle_extract.argtypes = [P, I, I, P, I, D, I, D, P]
        l.oracle_synthesize.argtypes = [P, P, P, I, I, I, D, P, I]
        l.oracle_decode_aperiodicity.argtypes = [P, I, I, I, P]
        l.oracle_decode_aperiodicity.restype = None
        l.oracle_synthesize_features.argtypes = [P, I, I, I, D, P, I]
        l.oracle_cheaptrick_fft_size.argtypes = [I, D]
        l.oracle_num_aperiodicities.argtypes = [I]
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data


def cheaptrick(x, f0, t, fs, q1=-0.15, fft_size=None):
    l = lib()
    x = np.ascontiguousarray(x, np.float64)
    f0 = np.ascontiguousarray(f0, np.float64)
    t = np.ascontiguousarray(t, np.float64)
    if fft_size is None:
        fft_size = l.oracle_cheaptrick_fft_size(int(fs), 71.0)
    sp = np.empty((len(f0), fft_size // 2 + 1))
    l.oracle_cheaptrick(_p(x), len(x), int(fs), _p(f0), _p(t), len(f0), fft_size, float(q1), _p(sp))
    return sp


def d4c_coarse(x, f0, t, fs, threshold=0.85):
    l = lib()
    x = np.ascontiguousarray(x, np.float64)
    f0 = np.ascontiguousarray(f0, np.float64)
    t = np.ascontiguousarray(t, np.float64)
    nap = l.oracle_num_aperiodicities(int(fs))
    coarse = np.zeros((len(f0), nap))
    voiced = np.zeros(len(f0), np.uint8)
    l.oracle_d4c_coarse(_p(x), len(x), int(fs), _p(f0), _p(t), len(f0), float(threshold), _p(coarse), _p(voiced))
    return voiced.astype(bool), coarse


def bap_from_coarse(coarse, voiced, fs, fft_size):
    l = lib()
    coarse = np.ascontiguousarray(coarse, np.float64)
    v = np.ascontiguousarray(voiced, np.uint8)
    bap = np.empty_like(coarse)
    l.oracle_bap_from_coarse(_p(coarse), _p(v), len(v), int(fs), int(fft_size), _p(bap))
    return bap


def mcep(spec, order, alpha, is_power=False, miniter=2, maxiter=30, threshold=0.001, eps=1e-8):
    l = lib()
    spec = np.ascontiguousarray(spec, np.float64)
    T, K = spec.shape
    mc = np.empty((T, order + 1))
    iters = np.zeros(T, np.int32)
    rc = l.oracle_mcep(_p(spec), 1 if is_power else 0, T, 2 * (K - 1), int(order), float(alpha), miniter, maxiter, float(threshold),
                       float(eps), _p(mc), _p(iters))
    if rc:
        raise RuntimeError("oracle_mcep failed with code %d" % rc)
    return mc, iters


def lf0_vuv(f0, thr=30.0, lf0_zero=0.0):
    l = lib()
    f0 = np.ascontiguousarray(f0, np.float64)
    lf0 = np.empty(len(f0), np.float32)
    vuv = np.empty(len(f0), np.float32)
    l.oracle_lf0_vuv(_p(f0), len(f0), float(thr), float(lf0_zero), _p(lf0), _p(vuv))
    return lf0, vuv


def extract(wave_i16, fs, f0, num_coded_sps, alpha, preemphasis=0.0):
    """One utterance: int16 wave + cached F0 -> [T, D + 2 + nap] float32 (releases the GIL: thread-pool friendly)."""
    l = lib()
    w = np.ascontiguousarray(wave_i16, np.int16)
    f0 = np.ascontiguousarray(f0, np.float64)
    nap = l.oracle_num_aperiodicities(int(fs))
    feats = np.empty((len(f0), num_coded_sps + 2 + nap), np.float32)
    rc = l.oracle_extract(_p(w), len(w), int(fs), _p(f0), len(f0), float(preemphasis), int(num_coded_sps), float(alpha), _p(feats))
    if rc:
        raise RuntimeError("oracle_extract failed with code %d" % rc)
    return feats


def synthesize(f0, sp, ap, fs, frame_period=5.0):
    """pyworld.synthesize on the CPU oracle (C port of oracle/world_np.py::synthesize)."""
    l = lib()
    f0 = np.ascontiguousarray(f0, np.float64)
    sp = np.ascontiguousarray(sp, np.float64)
    ap = np.ascontiguousarray(ap, np.float64)
    T = len(f0)
    y = np.zeros(int(T * frame_period * fs / 1000))
    l.oracle_synthesize(_p(f0), _p(sp), _p(ap), T, 2 * (sp.shape[1] - 1), int(fs), float(frame_period), _p(y), len(y))
    return y


def decode_aperiodicity(bap, fs, fft_size):
    l = lib()
    bap = np.ascontiguousarray(bap, np.float64)
    ap = np.empty((bap.shape[0], fft_size // 2 + 1))
    l.oracle_decode_aperiodicity(_p(bap), bap.shape[0], int(fs), int(fft_size), _p(ap))
    return ap


def synthesize_features(feats, fs, num_coded_sps, alpha):
    """One utterance of Synthesiser.run_world_synth: [T, D + 2 + nap] float32 rows -> float32 waveform (releases the GIL)."""
    l = lib()
    feats = np.ascontiguousarray(feats, np.float32)
    T = feats.shape[0]
    y = np.empty(int(T * 5.0 * fs / 1000), np.float32)
    n = l.oracle_synthesize_features(_p(feats), T, int(fs), int(num_coded_sps), float(alpha), _p(y), len(y))
    if n < 0:
        raise RuntimeError("oracle_synthesize_features failed with code %d" % n)
    return y[:n]
