"""ctypes binding of libb200world.so (include/b200world.h).  No torch types cross this boundary: only raw device
pointers, sizes and a cudaStream_t.  There is NO CPU fallback: if the library is missing and cannot be built, or a
call fails, the error is raised."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2W_LIB selects another build of the same ABI (kernel-variant experiments, scripts/gpu_kbench.py)
LIB_PATH = os.environ.get("B2W_LIB") or os.path.join(_HERE, "libb200world.so")

B2W_F64, B2W_F32, B2W_I16 = 0, 1, 2
STATUS_F0_TOO_HIGH = 1
STATUS_ZERO_PERIODOGRAM = 2
STATUS_SOLVE_FAILED = 4
STATUS_NOT_CONVERGED = 8

c_void_p, c_int32, c_int64, c_double = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double


class Batch(ctypes.Structure):
    """struct b2w_batch"""
    _fields_ = [("x", c_void_p), ("x_dtype", c_int32), ("num_utts", c_int32), ("preemphasis", c_double),
                ("utt_sample_offset", c_void_p), ("frame_utt", c_void_p), ("f0", c_void_p), ("t", c_void_p),
                ("num_frames", c_int64), ("fs", c_int32), ("reserved", c_int32)]


_SIGNATURES = {
    "b2w_version": (c_int32, []),
    "b2w_last_error": (ctypes.c_char_p, []),
    "b2w_cheaptrick": (c_int32, [ctypes.POINTER(Batch), c_int32, c_double, c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
    "b2w_d4c_coarse": (c_int32, [ctypes.POINTER(Batch), c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2w_d4c_coarse_f64": (c_int32, [ctypes.POINTER(Batch), c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2w_d4c_expand": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "b2w_bap_from_coarse": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "b2w_code_aperiodicity": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "b2w_decode_aperiodicity": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "b2w_decode_aperiodicity_f32": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "b2w_mcep_pad": (c_int32, [c_int32]),
    "b2w_mcep_tables_host": (c_int32, [c_int32, c_double, c_int32, c_void_p, c_void_p, c_void_p]),
    "b2w_mcep": (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_int32, c_int32, c_double, c_int32, c_int32, c_double,
                           c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p,
                           c_void_p]),
    "b2w_mcep_tc_stream_floats": (c_int64, [c_int32]),
    "b2w_mcep_tc_pretile": (c_int32, [c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2w_mcep_tc": (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_int64, c_int32, c_int32, c_double, c_int32, c_int32, c_double,
                              c_double, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    "b2w_mc2sp": (c_int32, [c_void_p, c_int32, c_int64, c_int64, c_int32, c_int32, c_void_p, c_double, c_int32, c_void_p,
                            c_int32, c_void_p]),
    "b2w_mc2sp_tc": (c_int32, [c_void_p, c_int32, c_int64, c_int64, c_int32, c_int32, c_void_p, c_double, c_int32, c_void_p, c_void_p]),
    "b2w_lf0_vuv": (c_int32, [c_void_p, c_void_p, c_int32, c_double, c_double, c_void_p, c_void_p, c_int64, c_void_p]),
    "b2w_deltas": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_int64,
                             c_void_p]),
    "b2w_stats_accumulate": (c_int32, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    "b2w_synth_max_pulses": (c_int64, [c_int64, c_int32]),
    "b2w_synth_randn_table": (c_int32, [c_void_p, c_int64, c_void_p]),
    "b2w_synth_timebase_chunks": (c_int64, [c_int64]),
    "b2w_synth_timebase": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_double, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2w_synth_render": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int64, c_int32, c_double, c_int32, c_int64, c_void_p, c_void_p]),
    "b2w_synth_overlap_add": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int64,
                                        c_double, c_void_p, c_int32, c_void_p]),
    "b2w_synth_render_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int64, c_int32, c_double, c_int32, c_int64, c_void_p, c_void_p]),
    "b2w_synth_overlap_add_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int64,
                                            c_double, c_void_p, c_int32, c_void_p]),
    "b2w_allpass_forward": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    "b2w_allpass_forward_tc": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    "b2w_allpass_forward_masked": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p]),
    "b2w_allpass_backward_tc": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2w_allpass_backward_masked": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2w_allpass_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2w_mlpg_workspace_doubles": (c_int64, [c_int64, c_int32]),
    "b2w_mlpg": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "b2w_world_metrics": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "b2w_pad_normalise": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                    c_void_p, c_void_p]),
    "b2w_unpad_denormalise": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                        c_int32, c_void_p, c_int64, c_void_p]),
    "b2w_dio_num_bands": (c_int32, [c_double, c_double, c_double]),
    "b2w_dio_workspace_bytes": (c_int64, [c_int64, c_int32, c_int64, c_int32, c_double, c_double, c_double]),
    "b2w_dio": (c_int32, [ctypes.POINTER(Batch), c_int64, c_void_p, c_double, c_double, c_double, c_double, c_double, c_int32,
                          c_void_p, c_void_p, c_void_p]),
    "b2w_stonemask": (c_int32, [ctypes.POINTER(Batch), c_void_p, c_void_p]),
    "b2w_mgcep": (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_int32, c_int32, c_double, c_int32, c_int32, c_double, c_double,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p,
                            c_void_p]),
    "b2w_mgc2sp": (c_int32, [c_void_p, c_int32, c_int64, c_int64, c_int32, c_int32, c_double, c_void_p, c_void_p, c_void_p, c_int32,
                             c_void_p]),
    "b2w_probe_fp64_fma": (c_int64, [c_int32, c_void_p, c_void_p]),
    "b2w_mcep_prof_read": (c_int32, [c_void_p]),
    "b2w_vtf_prof_read": (c_int32, [c_void_p]),
    "b2w_wav_probe": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32]),
    "b2w_wav_read_i16": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32]),
    "b2w_wav_write_pcm16": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32]),
    "b2w_npz_write_f32": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32]),
    "b2w_npz_probe": (c_int32, [c_void_p, c_int32, ctypes.c_char_p, c_void_p, c_void_p, c_int32]),
    "b2w_npz_read_f32": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                   c_int32]),
    "b2w_crc32": (ctypes.c_uint32, [c_void_p, c_int64, c_int32]),
    "b2w_cheaptrick_fft_size": (c_int32, [c_int32, c_double]),
    "b2w_num_aperiodicities": (c_int32, [c_int32]),
    "b2w_d4c_fft_size": (c_int32, [c_int32]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))

_lock = threading.Lock()
_lib = None


class B200WorldError(RuntimeError):
    pass


def load():
    """Loads (building it first if the .so is absent) and returns the ctypes library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build  # nvcc must exist; raises otherwise
            _build.build(verbose=False)
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise B200WorldError("cannot load %s: %s (run `python -m idiaptts_b200.build`)" % (LIB_PATH, e)) from e
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        if lib.b2w_version() != 1:
            raise B200WorldError("libb200world.so version %d, binding expects 1" % lib.b2w_version())
        _lib = lib
        return lib


def check(rc, what):
    if rc != 0:
        msg = load().b2w_last_error().decode("utf-8", "replace")
        if rc < 0:
            raise ValueError("%s: %s" % (what, msg))
        raise B200WorldError("%s: CUDA error %d: %s" % (what, rc, msg))
