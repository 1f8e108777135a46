"""Drop-in for the pysptk functions IdiapTTS calls on the WORLD feature path (numpy float64 in / out), executed on the
GPU by libb200world.so.

Reference call sites: idiaptts/src/data_preparation/audio/AudioProcessing.py:146 (pysptk.mcep), :252 (pysptk.mgc2sp,
gamma = 0), :40 (pysptk.util.mcepalpha).  Like pysptk, 1-D input is one frame and 2-D input is frames on axis 0."""
import numpy as np
import torch

from .. import ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("idiaptts_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as2d(x):
    x = np.ascontiguousarray(x, np.float64)
    if x.ndim == 1:
        return x[None, :], True
    if x.ndim != 2:
        raise ValueError("expected a 1-D frame or a 2-D [frames, bins] array")
    return x, False


def mcep(x, order=25, alpha=0.35, miniter=2, maxiter=30, threshold=0.001, etype=0, eps=0.0, min_det=1.0e-6, itype=0):
    """pysptk.mcep for spectral input: itype 3 (amplitude) or 4 (periodogram), etype 0/1 (eps floor)."""
    if itype not in (3, 4):
        raise NotImplementedError("only itype=3 (amplitude) and itype=4 (periodogram) inputs are accelerated "
                                  "(IdiapTTS uses itype=3, AudioProcessing.py:146)")
    if etype not in (0, 1):
        raise ValueError("etype must be 0 or 1 on this path")
    if etype == 0 and eps != 0.0:
        raise ValueError("eps cannot be specified for etype = 0")
    if etype == 1 and eps < 0.0:
        raise ValueError("eps: negative value is not allowed")
    x2, single = _as2d(x)
    plane = torch.from_numpy(x2).to(_device())
    mc, status = ops.mcep(plane, order, alpha, is_power=(itype == 4), miniter=miniter, maxiter=maxiter, threshold=threshold,
                          eps=eps if etype == 1 else 0.0, out_dtype=torch.float64)
    out = mc.cpu().numpy()
    ops.raise_for_status(status, "mcep")
    return out[0] if single else out


def mgcep(x, order=25, alpha=0.35, gamma=0.0, num_recursions=None, miniter=2, maxiter=30, threshold=0.001, etype=0, eps=0.0,
          min_det=1.0e-6, itype=0, otype=0):
    """pysptk.mgcep for spectral input (itype 3 amplitude / 4 periodogram, etype 0/1, otype 0): mel-generalised cepstral analysis
    (SURVEY 8f N3).  PARITY UNPINNED: restated from the published criterion (csrc/mgcep.cu), not from SPTK's source."""
    if itype not in (3, 4):
        raise NotImplementedError("only itype=3 (amplitude) and itype=4 (periodogram) inputs are accelerated "
                                  "(IdiapTTS uses itype=3, AudioProcessing.py:138)")
    if otype != 0:
        raise NotImplementedError("only otype=0 (cepstral coefficients) is implemented")
    if etype not in (0, 1):
        raise ValueError("etype must be 0 or 1 on this path")
    if not (-1.0 <= gamma <= 0.0):
        raise ValueError("gamma must be in [-1, 0]")
    x2, single = _as2d(x)
    plane = torch.from_numpy(x2).to(_device())
    mgc, status = ops.mgcep(plane, order, alpha, gamma, is_power=(itype == 4), miniter=miniter, maxiter=maxiter, threshold=threshold,
                            eps=eps if etype == 1 else 0.0, out_dtype=torch.float64)
    out = mgc.cpu().numpy()
    ops.raise_for_status(status, "mgcep")
    return out[0] if single else out


def mgc2sp(ceps, alpha=0.0, gamma=0.0, fftlen=256):
    """pysptk.mgc2sp.  Returns a complex array like pysptk; only the real part (log amplitude) is computed on this path because
    that is all the reference reads (AudioProcessing.py:256, :275 `amp_sp.real`); the imaginary part is 0.  gamma != 0 evaluates
    log |1 + gamma C|^(1/gamma) directly (pysptk goes through a 512-term cepstrum: 3e-8 relative apart, oracle/mgc_np.py)."""
    c2, single = _as2d(ceps)
    mc = torch.from_numpy(c2).to(_device())
    if gamma != 0.0:
        sp = torch.log(ops.mgc2sp(mc, alpha, gamma, fftlen, out_dtype=torch.float64)).cpu().numpy()
    else:
        sp = ops.mc2sp(mc, alpha, fftlen, scale=1.0, do_exp=False, out_dtype=torch.float64).cpu().numpy()
    out = sp.astype(np.complex128)
    return out[0] if single else out


def mc2sp(mc, alpha, fftlen):
    """pysptk.mc2sp: mel-cepstrum -> power spectrum."""
    c2, single = _as2d(mc)
    t = torch.from_numpy(c2).to(_device())
    sp = ops.mc2sp(t, alpha, fftlen, scale=2.0, do_exp=True, out_dtype=torch.float64).cpu().numpy()
    return sp[0] if single else sp


def mcepalpha(fs, start=0.0, stop=1.0, step=0.001, num_points=1000):
    """pysptk.util.mcepalpha (host scalar search, fp64): the all-pass constant approximating the mel scale at fs."""
    def melscale(fs_, n):
        v = np.log(1 + (fs_ / 2.0) / n * np.arange(0, n) / 1000.0)
        return v / v[-1]

    def warp(a, n):
        omega = np.pi / n * np.arange(0, n)
        w = np.arctan((1 - a * a) * np.sin(omega) / ((1 + a * a) * np.cos(omega) - 2 * a))
        w[w < 0] += np.pi
        return w / w[-1]

    cands = np.arange(start, stop, step)
    mel = melscale(fs, num_points)
    dist = [np.sqrt(np.mean((mel - warp(a, num_points)) ** 2)) for a in cands]
    return cands[int(np.argmin(dist))]


class util:  # pysptk.util.mcepalpha
    mcepalpha = staticmethod(mcepalpha)
