"""The WORLD/SPTK part of the reference's idiaptts/src/data_preparation/audio/AudioProcessing.py with the same static
interface, executed on the GPU: fs_to_mgc_alpha :33-40, fs_to_frame_length :53-60, fs_to_num_bap :70-71, get_raw :108-120,
extract_mcep :143-153, mcep_to_amp_sp :248-256, decode_sp :304-327, depreemphasis :330-331.

Out of scope: librosa STFT / mel filter banks / Griffin-Lim, generalised cepstra (sp_type "mgc")
and Merlin post-filtering; they raise NotImplementedError instead of silently doing something else."""
import os
import wave

import logging

import numpy as np
import torch

from . import ops
from .compat import pysptk as _sptk


class AudioProcessing:
    mgc_gamma = -1. / 3.

    @staticmethod
    def fs_to_mgc_alpha(fs):
        return _sptk.mcepalpha(fs)

    @staticmethod
    def fs_to_frame_length(fs):
        return ops.get_cheaptrick_fft_size(fs)

    @staticmethod
    def fs_to_num_bap(fs):
        return ops.get_num_aperiodicities(fs)

    @staticmethod
    def read_wav(audio_name):
        """PCM wav -> (samples, fs).  int16 files come back as int16 (the kernels scale by 1/32768 on load, which is what
        soundfile.read returns as float64); other widths are converted to float64 in [-1, 1)."""
        with wave.open(audio_name, "rb") as w:
            fs, width, ch, n = w.getframerate(), w.getsampwidth(), w.getnchannels(), w.getnframes()
            data = w.readframes(n)
        if width == 2:
            x = np.frombuffer(data, dtype=np.int16)
        elif width == 4:
            x = np.frombuffer(data, dtype=np.int32).astype(np.float64) / 2147483648.0
        elif width == 1:
            x = (np.frombuffer(data, dtype=np.uint8).astype(np.float64) - 128.0) / 128.0
        else:
            raise ValueError("unsupported sample width %d in %s" % (width, audio_name))
        if ch > 1:
            raise ValueError("%s: only mono files are supported on the WORLD path" % audio_name)
        return x, fs

    @staticmethod
    def get_raw(audio_name, preemphasis=0.0):
        """Raw audio in [-1, 1) as float64 with pre-emphasis applied (reference semantics, host side)."""
        x, fs = AudioProcessing.read_wav(audio_name)
        raw = x.astype(np.float64) / 32768.0 if x.dtype == np.int16 else x
        raw = np.append(raw[0], raw[1:] - preemphasis * raw[:-1])
        return raw, fs

    @staticmethod
    def extract_mcep(amp_sp, num_coded_sps, mgc_alpha):
        """pysptk.mcep(amp_sp, order, alpha, eps=1e-8, etype=1, itype=3) on the GPU -> float32 [T, num_coded_sps]."""
        mcep = _sptk.mcep(amp_sp, order=num_coded_sps - 1, alpha=mgc_alpha, eps=1.0e-8, min_det=0.0, etype=1, itype=3)
        return mcep.astype(np.float32, copy=False)

    @staticmethod
    def extract_mgc(amp_sp, fs=None, num_coded_sps=60, mgc_alpha=None):
        """pysptk.mgcep(amp_sp, order, alpha, gamma = -1/3, eps=1e-8, etype=1, itype=3) on the GPU -> float32 [T, num_coded_sps]
        (reference :123-140; SURVEY 8f N3, parity unpinned)."""
        if mgc_alpha is None:
            assert fs is not None, "Either sampling rate or mgc alpha has to be given."
            mgc_alpha = AudioProcessing.fs_to_mgc_alpha(fs)
        mgc = _sptk.mgcep(amp_sp, order=num_coded_sps - 1, alpha=mgc_alpha, gamma=AudioProcessing.mgc_gamma, eps=1.0e-8, min_det=0.0,
                          etype=1, itype=3)
        return mgc.astype(np.float32, copy=False)

    @staticmethod
    def mcep_to_amp_sp(mcep, fs, alpha=None):
        if alpha is None:
            alpha = AudioProcessing.fs_to_mgc_alpha(fs)
        if not torch.cuda.is_available():
            raise RuntimeError("idiaptts_b200 needs a CUDA device; there is no CPU fallback")
        mc = torch.from_numpy(np.ascontiguousarray(mcep, dtype=np.float64)).cuda()
        amp = ops.mc2sp(mc, alpha, AudioProcessing.fs_to_frame_length(fs), scale=1.0, do_exp=True, out_dtype=torch.float32)
        return amp.cpu().numpy()

    @staticmethod
    def mgc_to_amp_sp(mgc, fs, alpha=None, gamma=None, n_fft=None):
        """exp(Re pysptk.mgc2sp(mgc, alpha, gamma, n_fft)) as float32 (reference :259-275)."""
        if alpha is None:
            alpha = AudioProcessing.fs_to_mgc_alpha(fs)
        if gamma is None:
            gamma = AudioProcessing.mgc_gamma
        if n_fft is None:
            n_fft = AudioProcessing.fs_to_frame_length(fs)
        if not torch.cuda.is_available():
            raise RuntimeError("idiaptts_b200 needs a CUDA device; there is no CPU fallback")
        c = torch.from_numpy(np.ascontiguousarray(mgc, dtype=np.float64)).cuda()
        return ops.mgc2sp(c, alpha, gamma, n_fft, out_dtype=torch.float32).cpu().numpy()

    @staticmethod
    def merlin_post_filter(coded_sp, alpha, fft_size=1024):
        """nnmnkwii.postfilters.merlin_post_filter (reference :308-311) on the GPU."""
        if not torch.cuda.is_available():
            raise RuntimeError("idiaptts_b200 needs a CUDA device; there is no CPU fallback")
        c = torch.from_numpy(np.ascontiguousarray(coded_sp, dtype=np.float64)).cuda()
        return ops.merlin_post_filter(c, alpha, fft_size).cpu().numpy()

    @staticmethod
    def decode_sp(coded_sp, sp_type="mcep", fs=None, alpha=None, mgc_gamma=None, n_fft=None, post_filtering=False):
        if post_filtering:
            if sp_type in ["mcep", "mgc"]:
                coded_sp = AudioProcessing.merlin_post_filter(coded_sp, AudioProcessing.fs_to_mgc_alpha(fs))
            else:
                logging.warning("Post-filtering only implemented for cepstrum features.")
        if sp_type == "mcep":
            return AudioProcessing.mcep_to_amp_sp(coded_sp, fs, alpha)
        if sp_type == "mgc":
            return AudioProcessing.mgc_to_amp_sp(coded_sp, fs, alpha, mgc_gamma, n_fft)
        if sp_type == "amp_sp":
            return coded_sp
        raise NotImplementedError("Unknown or unsupported feature type {}. No decoding method available.".format(sp_type))

    @staticmethod
    def depreemphasis(raw, preemphasis):
        """lfilter([1], [1, -p], raw): y[n] = x[n] + p y[n-1] (float64).  Identity copy for p == 0."""
        raw = np.asarray(raw)
        out = raw.astype(np.float64)
        if preemphasis != 0.0:
            prev = 0.0
            for i in range(len(out)):
                prev = out[i] + preemphasis * prev
                out[i] = prev
        return out
