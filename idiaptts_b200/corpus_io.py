"""Shard-at-a-time file IO around the GPU batch: wav files into the packed int16 sample buffer, packed float32
feature matrices out to / in from per-utterance .npz archives.  Thin wrappers over the host entry points of libb200world.so
(csrc/corpus_io.cu: a pool of threads per call); the formats are numpy's and soundfile's, so files written here load with
numpy.load and the readers take what numpy.savez wrote.

Replaces, for a whole shard per call, the per-utterance loops of the reference:
    soundfile.read in AudioProcessing.get_raw                idiaptts/src/data_preparation/audio/AudioProcessing.py:108-120
    numpy.savez in LabelGen.save_output / WorldFeatLabelGen.save_output   LabelGen.py:63-101, world/WorldFeatLabelGen.py:1121-1172
    numpy.load in WorldFeatLabelGen.load_sample              world/WorldFeatLabelGen.py:459-567
"""
import ctypes

import numpy as np

from . import _lib

DEFAULT_THREADS = 0  # one per hardware thread (at most 64)


def _paths(paths):
    enc = [p.encode() if isinstance(p, str) else bytes(p) for p in paths]
    return (ctypes.c_char_p * len(enc))(*enc), enc


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()


def probe_wavs(paths, threads=DEFAULT_THREADS):
    """RIFF headers of many files -> dict of arrays: num_samples (int64), fs, bits, channels (int32), data_offset (int64)."""
    n = len(paths)
    out = dict(num_samples=np.zeros(n, np.int64), fs=np.zeros(n, np.int32), bits=np.zeros(n, np.int32),
               channels=np.zeros(n, np.int32), data_offset=np.zeros(n, np.int64))
    arr, _keep = _paths(paths)
    _lib.check(_lib.load().b2w_wav_probe(arr, n, _ptr(out["num_samples"]), _ptr(out["fs"]), _ptr(out["bits"]), _ptr(out["channels"]),
                                         _ptr(out["data_offset"]), threads), "probe_wavs")
    return out


def read_wavs_i16(paths, info=None, pin=False, threads=DEFAULT_THREADS):
    """16-bit mono PCM files -> (samples, sample_off, fs): ONE packed int16 torch tensor (pinned on request: it goes to the device
    as the b2w_batch waveform) and the int64 offsets [U + 1].  Raises ValueError when a file is not 16-bit mono or the sampling
    rates differ: the caller takes the general per-file path then."""
    import torch
    info = info if info is not None else probe_wavs(paths, threads)
    n = len(paths)
    if n and (np.any(info["bits"] != 16) or np.any(info["channels"] != 1)):
        bad = int(np.flatnonzero((info["bits"] != 16) | (info["channels"] != 1))[0])
        raise ValueError("%s: %d-bit, %d channel(s); the packed reader takes 16-bit mono" % (paths[bad], info["bits"][bad], info["channels"][bad]))
    if n and np.any(info["fs"] != info["fs"][0]):
        bad = int(np.flatnonzero(info["fs"] != info["fs"][0])[0])
        raise ValueError("mixed sampling rates in one shard ({} vs {})".format(int(info["fs"][0]), int(info["fs"][bad])))
    sample_off = np.concatenate(([0], np.cumsum(info["num_samples"]))).astype(np.int64)
    samples = torch.empty(int(sample_off[-1]), dtype=torch.int16, pin_memory=bool(pin and sample_off[-1] > 0))
    arr, _keep = _paths(paths)
    _lib.check(_lib.load().b2w_wav_read_i16(arr, n, _ptr(info["data_offset"]), _ptr(sample_off), samples.data_ptr(), threads),
               "read_wavs_i16")
    return samples, sample_off, (int(info["fs"][0]) if n else 0)


def _read_into(paths, data_offset, sample_off, samples, threads=DEFAULT_THREADS):
    """read_wavs_i16 into a caller-owned int16 tensor (gen_data's reusable pinned buffers); the caller has probed the files."""
    assert str(samples.dtype) == "torch.int16" and samples.numel() >= int(sample_off[-1])
    arr, _keep = _paths(paths)
    _lib.check(_lib.load().b2w_wav_read_i16(arr, len(paths), _ptr(np.ascontiguousarray(data_offset, np.int64)),
                                            _ptr(np.ascontiguousarray(sample_off, np.int64)), samples.data_ptr(), threads), "read_wavs_i16")


def write_wavs_pcm16(paths, samples, sample_off, fs, threads=DEFAULT_THREADS):
    """Packed float32 waveforms (numpy or CPU torch, utterance u at sample_off[u]:sample_off[u + 1]) -> one 16-bit mono PCM wav per
    utterance, clip(round(32767 x)) as Synthesiser.write_wav does."""
    assert samples.dtype in (np.float32,) or str(samples.dtype) == "torch.float32", "float32 samples"
    sample_off = np.ascontiguousarray(sample_off, np.int64)
    assert len(sample_off) == len(paths) + 1 and int(sample_off[-1]) <= int(samples.shape[0])
    arr, _keep = _paths(paths)
    _lib.check(_lib.load().b2w_wav_write_pcm16(arr, len(paths), _ptr(sample_off), _ptr(samples), int(fs), threads), "write_wavs_pcm16")


def _key_args(keys, col_offset, cols):
    enc = [k.encode() for k in keys]
    karr = (ctypes.c_char_p * len(enc))(*enc)
    return karr, enc, np.ascontiguousarray(col_offset, np.int32), np.ascontiguousarray(cols, np.int32)


def write_npz(paths, keys, col_offset, cols, frame_off, feats, threads=DEFAULT_THREADS):
    """For every utterance u: paths[u] <- archive {keys[k]: feats[frame_off[u]:frame_off[u + 1], col_offset[k]:col_offset[k] + cols[k]]}.
    feats: C-contiguous float32 [F, W] host matrix (numpy array or CPU torch tensor).  paths carry the .npz suffix."""
    assert feats.dtype in (np.float32,) or str(feats.dtype) == "torch.float32", "float32 features"
    W = int(feats.shape[1])
    frame_off = np.ascontiguousarray(frame_off, np.int64)
    assert len(frame_off) == len(paths) + 1 and int(frame_off[-1]) <= int(feats.shape[0])
    karr, _k, co, cc = _key_args(keys, col_offset, cols)
    arr, _keep = _paths(paths)
    _lib.check(_lib.load().b2w_npz_write_f32(arr, len(paths), karr, len(keys), _ptr(co), _ptr(cc), _ptr(frame_off), _ptr(feats), W, threads),
               "write_npz")


def probe_npz(paths, key, threads=DEFAULT_THREADS):
    """Shape of array `key` in every archive -> (rows int64 [U], cols int32 [U])."""
    n = len(paths)
    rows, cols = np.zeros(n, np.int64), np.zeros(n, np.int32)
    arr, _keep = _paths(paths)
    _lib.check(_lib.load().b2w_npz_probe(arr, n, key.encode(), _ptr(rows), _ptr(cols), threads), "probe_npz")
    return rows, cols


def read_npz(paths, keys, col_offset, cols, frame_off, feats, verify_crc=True, threads=DEFAULT_THREADS):
    """Inverse of write_npz: fills the column blocks of the packed float32 matrix `feats` [F, W] (numpy or CPU torch, e.g. pinned)
    from the archives; every array must have exactly frame_off[u + 1] - frame_off[u] rows and cols[k] columns."""
    assert feats.dtype in (np.float32,) or str(feats.dtype) == "torch.float32", "float32 features"
    W = int(feats.shape[1])
    frame_off = np.ascontiguousarray(frame_off, np.int64)
    assert len(frame_off) == len(paths) + 1 and int(frame_off[-1]) <= int(feats.shape[0])
    karr, _k, co, cc = _key_args(keys, col_offset, cols)
    arr, _keep = _paths(paths)
    _lib.check(_lib.load().b2w_npz_read_f32(arr, len(paths), karr, len(keys), _ptr(co), _ptr(cc), _ptr(frame_off), _ptr(feats), W,
                                            1 if verify_crc else 0, threads), "read_npz")
