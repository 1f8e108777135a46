"""The C-ABI shared library loads on a machine without a GPU and exports exactly what include/b200world.h declares;
host-side (no-GPU) entry points behave; device entry points reject bad arguments without launching anything."""
import ctypes
import os
import re

import numpy as np
import pytest

from idiaptts_b200 import _lib
from oracle import sptk_np, world_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200world.h")).read()
    return sorted(set(re.findall(r"B2W_API\s+[\w\s\*]+?\b(b2w_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    # and the binding knows every one of them
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    # nm view: nothing else with the b2w_ prefix is exported
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if " T " in line and "b2w_" in line)
    assert exported == declared


def test_scalar_helpers_match_world():
    lib = _lib.load()
    assert lib.b2w_version() == 1
    for fs in (8000, 16000, 22050, 24000, 32000, 44100, 48000):
        assert lib.b2w_cheaptrick_fft_size(fs, 71.0) == world_np.get_cheaptrick_fft_size(fs)
        assert lib.b2w_num_aperiodicities(fs) == world_np.get_num_aperiodicities(fs)
        assert lib.b2w_d4c_fft_size(fs) == world_np.get_d4c_fft_size(fs)
    assert lib.b2w_synth_max_pulses(16000, 16000) >= 1200
    assert [lib.b2w_mcep_pad(n) for n in (1, 4, 61, 119)] == [4, 4, 64, 120]


@pytest.mark.parametrize("order,alpha,fft", [(19, 0.58, 1024), (59, 0.455, 1024), (24, 0.42, 512)])
def test_mcep_tables_equal_composed_allpass_maps(order, alpha, fft):
    """b2w_mcep_tables_host == (freqt / frqtr recursions of the oracle) composed with the real (I)DFT, in fp64."""
    lib = _lib.load()
    K = fft // 2 + 1
    m = order
    np0, np2, mp = lib.b2w_mcep_pad(m + 2), lib.b2w_mcep_pad(2 * m + 1), lib.b2w_mcep_pad(m + 1)
    m0t, cmat, m2t = np.empty((K, np0)), np.empty((mp, K)), np.empty((K, np2))
    assert lib.b2w_mcep_tables_host(m, alpha, fft, m0t.ctypes.data, cmat.ctypes.data, m2t.ctypes.data) == 0
    f2 = fft // 2
    A = sptk_np.freqt_matrix(f2, m, alpha)
    B = sptk_np.freqt_matrix(m, f2, -alpha)
    R = sptk_np.frqtr_matrix(f2, 2 * m, alpha)
    n, j = np.arange(K)[:, None], np.arange(K)[None, :]
    cosm = np.cos(2 * np.pi * n * j / fft)
    w = np.full(K, 2.0)
    w[0] = w[-1] = 1.0
    F = cosm * w[None, :] / fft
    D = np.ones(K)
    D[0] = D[-1] = 0.5
    np.testing.assert_allclose(m0t[:, :m + 1], (A @ (D[:, None] * F)).T, atol=1e-13)
    np.testing.assert_allclose(m0t[:, m + 1], (D[:, None] * F)[0], atol=1e-15)
    np.testing.assert_allclose(cmat[:m + 1], (cosm @ B).T, atol=1e-12)
    assert np.all(cmat[m + 1:] == 0) and np.all(m0t[:, m + 2:] == 0) and np.all(m2t[:, 2 * m + 1:] == 0)
    np.testing.assert_allclose(m2t[:, :2 * m + 1], (R @ F).T, atol=1e-13)


def test_argument_errors_do_not_launch():
    lib = _lib.load()
    st = ctypes.c_int32(0)
    b = _lib.Batch()
    b.x_dtype = 7
    rc = lib.b2w_cheaptrick(b, 1024, -0.15, 0, 0, 513, ctypes.addressof(st), 0)
    assert rc < 0 and b"b2w_cheaptrick" in lib.b2w_last_error()
    rc = lib.b2w_mcep(0, 0, 0, 10, 1024, 59, 0.5, 2, 30, 1e-3, 1e-8, 0, 0, 0, 0, 1, 60, 0, 0, 0)
    assert rc < 0 and b"null" in lib.b2w_last_error()
    rc = lib.b2w_mcep_tables_host(0, 0.5, 1000, 0, 0, 0)
    assert rc < 0
    with pytest.raises(ValueError):
        _lib.check(rc, "tables")


def test_operators_refuse_cpu_tensors():
    import torch
    from idiaptts_b200 import ops
    with pytest.raises(ValueError, match="CUDA"):
        ops.code_aperiodicity(torch.ones(3, 513, dtype=torch.float64), 16000)
    with pytest.raises(ValueError, match="CUDA"):
        ops.allpass_forward(torch.zeros(2, 60), torch.zeros(2), 60)
