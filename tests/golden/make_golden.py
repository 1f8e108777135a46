"""Builds tests/golden/ljspeech_world_golden.npz from the reference's own test fixtures.

Run HERE (the container that has /root/reference); the GPU box only sees the committed .npz.

Source fixtures (read-only, data not code):
  /root/reference/test/integration/fixtures/database/wav/LJ001-000{1..9}.wav      16 kHz mono int16
  /root/reference/test/integration/fixtures/WORLD/cmp_mcep20/LJ001-000{1..9}.cmp  float32 [T x 67]:
        [mcep20 d dd | lf0 d dd | vuv | bap d dd], produced by the reference pipeline with pre-emphasis 0.97,
        alpha 0.58
  /root/reference/test/integration/fixtures/WORLD/{mcep20,lf0,bap}/{stats,mean-std_dev}.bin
        legacy stats: int32 N + float64 [2 x d]  (MeanStdDevExtractor.py:131-137)
  /root/reference/test/integration/fixtures/WORLD/{lf0,vuv}/<id>.{lf0,vuv}  float32 (outputs of interpolate_lin for a
        different F0 run: usable as a known-answer test for the interpolation only)
"""
import os
import struct
import wave

import numpy as np

REF = "/root/reference/test/integration/fixtures"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ljspeech_world_golden.npz")


def read_legacy_bin(path):
    with open(path, "rb") as f:
        n = struct.unpack("i", f.read(4))[0]
        data = np.fromfile(f, dtype=np.float64).reshape(2, -1)
    return n, data


def main():
    out = {}
    ids = ["LJ001-%04d" % i for i in range(1, 10)]
    out["ids"] = np.array(ids)
    for id_ in ids:
        w = wave.open(os.path.join(REF, "database/wav", id_ + ".wav"))
        assert w.getnchannels() == 1 and w.getsampwidth() == 2 and w.getframerate() == 16000
        out[id_ + "/wav"] = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
        out[id_ + "/cmp"] = np.fromfile(os.path.join(REF, "WORLD/cmp_mcep20", id_ + ".cmp"), np.float32).reshape(-1, 67)
        out[id_ + "/lf0_other"] = np.fromfile(os.path.join(REF, "WORLD/lf0", id_ + ".lf0"), np.float32)
        out[id_ + "/vuv_other"] = np.fromfile(os.path.join(REF, "WORLD/vuv", id_ + ".vuv"), np.float32)
    for feat in ("mcep20", "lf0", "bap"):
        for kind in ("stats", "mean-std_dev"):
            n, data = read_legacy_bin(os.path.join(REF, "WORLD", feat, kind + ".bin"))
            out["stats/%s/%s/n" % (feat, kind)] = np.int64(n)
            out["stats/%s/%s/data" % (feat, kind)] = data
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
