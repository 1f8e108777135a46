"""idiaptts_b200: the WORLD vocoder feature hot path of IdiapTTS as hand-written CUDA for B200 (sm_100a).

    csrc/                    CUDA kernels + the C ABI (include/b200world.h) -> libb200world.so
    _lib.py, ops.py          ctypes binding; device-tensor operators
    pipeline.py              fused extraction / synthesis over ragged batches
    compat/pyworld.py, compat/pysptk.py   pyworld- / pysptk-compatible call signatures
    WorldFeatLabelGen.py, AudioProcessing.py, Synthesiser.py, MeanStdDevExtractor.py, MeanCovarianceExtractor.py,
    AllPassWarp.py           the reference's own entry points for this path (same names, arguments, error behaviour)
"""
__version__ = "0.1.0"
