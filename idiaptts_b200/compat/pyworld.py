"""Drop-in for the pyworld functions IdiapTTS calls on the WORLD feature path (numpy float64 in / out), executed on the
GPU by libb200world.so.

Reference call sites (idiaptts/src/...): data_preparation/world/WorldFeatLabelGen.py:792 (wav2world), :805
(code_aperiodicity), :940 (decode_aperiodicity), :943 (synthesize); data_preparation/audio/AudioProcessing.py:60
(get_cheaptrick_fft_size), :71 (get_num_aperiodicities); Synthesiser.py:47.

F0: `dio` (speed = 1) and `stonemask` run on the device too; `wav2world` estimates the F0 track exactly as
pyworld does unless a cached track is passed as `f0=` (north_star: cached F0)."""
import numpy as np
import torch

from .. import ops

default_frame_period = 5.0
default_f0_floor = 71.0
default_f0_ceil = 800.0

get_cheaptrick_fft_size = ops.get_cheaptrick_fft_size
get_num_aperiodicities = ops.get_num_aperiodicities


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("idiaptts_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _batch(x, f0, temporal_positions, fs):
    x = np.ascontiguousarray(x, np.float64)
    f0 = np.ascontiguousarray(f0, np.float64)
    t = np.ascontiguousarray(temporal_positions, np.float64)
    if x.ndim != 1 or f0.ndim != 1 or t.ndim != 1:
        raise ValueError("x, f0 and temporal_positions must be 1-D")
    if len(f0) != len(t):  # pyworld: "Mismatched number of frames between F0 (..) and temporal positions (..)"
        raise ValueError("Mismatched number of frames between F0 ({:d}) and temporal positions ({:d}).".format(len(f0), len(t)))
    return ops.RaggedBatch.from_host([x], [f0], fs, ts=[t], device=_device())


def cheaptrick(x, f0, temporal_positions, fs, q1=-0.15, f0_floor=default_f0_floor, fft_size=None):
    """pyworld.cheaptrick -> spectrogram [T, fft_size/2+1] float64 (power)."""
    if fft_size is None:
        fft_size = get_cheaptrick_fft_size(fs, f0_floor)
    batch = _batch(x, f0, temporal_positions, fs)
    sp, status = ops.cheaptrick(batch, fft_size=fft_size, q1=q1, out_dtype=torch.float64)
    out = sp.cpu().numpy()
    ops.raise_for_status(status, "cheaptrick")
    return out


def d4c(x, f0, temporal_positions, fs, threshold=0.85, fft_size=None):
    """pyworld.d4c -> aperiodicity [T, fft_size/2+1] float64."""
    if fft_size is None:
        fft_size = get_cheaptrick_fft_size(fs, default_f0_floor)
    batch = _batch(x, f0, temporal_positions, fs)
    coarse, voiced, status = ops.d4c_coarse(batch, threshold=threshold, precision="f64")
    ap = ops.d4c_expand(coarse, voiced, fs, fft_size).cpu().numpy()
    ops.raise_for_status(status, "d4c")
    return ap


def dio(x, fs, f0_floor=default_f0_floor, f0_ceil=default_f0_ceil, channels_in_octave=2.0, frame_period=default_frame_period,
        speed=1, allowed_range=0.1):
    """pyworld.dio -> (f0 [T], temporal_positions [T]) float64.  Only speed = 1 (pyworld's default, no decimation)."""
    if speed != 1:
        raise ValueError("only speed = 1 (pyworld's default) is implemented")
    x = np.ascontiguousarray(x, np.float64)
    if x.ndim != 1:
        raise ValueError("x must be 1-D")
    T = ops.num_frames(len(x), fs, frame_period)
    t = np.arange(T) * frame_period / 1000.0
    batch = ops.RaggedBatch.from_host([x], [np.zeros(T)], fs, ts=[t], device=_device())
    f0 = ops.dio(batch, f0_floor, f0_ceil, channels_in_octave, frame_period, allowed_range)
    return f0.cpu().numpy(), t


def stonemask(x, f0, temporal_positions, fs):
    """pyworld.stonemask -> refined f0 [T] float64."""
    batch = _batch(x, f0, temporal_positions, fs)
    return ops.stonemask(batch).cpu().numpy()


def wav2world(x, fs, fft_size=None, frame_period=default_frame_period, f0=None):
    """pyworld.wav2world -> f0, sp, ap.  f0 = None: DIO + StoneMask as pyworld; else the supplied (cached) track is used."""
    if f0 is None:
        _f0, t = dio(x, fs, frame_period=frame_period)
        f0 = stonemask(x, _f0, t, fs)
    f0 = np.ascontiguousarray(f0, np.float64)
    t = np.arange(len(f0)) * frame_period / 1000.0
    return f0, cheaptrick(x, f0, t, fs, fft_size=fft_size), d4c(x, f0, t, fs, fft_size=fft_size)


def code_aperiodicity(aperiodicity, fs):
    ap = torch.from_numpy(np.ascontiguousarray(aperiodicity, np.float64)).to(_device())
    return ops.code_aperiodicity(ap, fs).cpu().numpy()


def decode_aperiodicity(coded_aperiodicity, fs, fft_size):
    bap = torch.from_numpy(np.ascontiguousarray(coded_aperiodicity, np.float64)).to(_device())
    return ops.decode_aperiodicity(bap, fs, fft_size).cpu().numpy()


def synthesize(f0, spectrogram, aperiodicity, fs, frame_period=default_frame_period):
    """pyworld.synthesize -> waveform float64 [int(T * frame_period * fs / 1000)]."""
    f0 = np.ascontiguousarray(f0, np.float64)
    sp = np.ascontiguousarray(spectrogram, np.float64)
    ap = np.ascontiguousarray(aperiodicity, np.float64)
    if sp.shape != ap.shape or sp.shape[0] != f0.shape[0]:  # pyworld raises ValueError on mismatched frames / bins
        raise ValueError("Mismatched number of frames between F0 ({:d}), spectrogram ({:d}) and aperiodicty ({:d})".format(
            f0.shape[0], sp.shape[0], ap.shape[0]))
    dev = _device()
    frame_off = torch.tensor([0, len(f0)], dtype=torch.int64, device=dev)
    y, _, status = ops.synthesize(torch.from_numpy(f0).to(dev), torch.from_numpy(sp).to(dev), torch.from_numpy(ap).to(dev),
                                  frame_off, fs, frame_period)
    out = y.cpu().numpy()
    ops.raise_for_status(status, "synthesize")
    return out
