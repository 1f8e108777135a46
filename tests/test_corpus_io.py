"""Shard-at-a-time file IO (csrc/corpus_io.cu through idiaptts_b200/corpus_io.py): host only, so these run without
a GPU.  The checkers are the formats' own Python implementations -- numpy.savez / numpy.load (what LabelGen.save_output and
WorldFeatLabelGen.load_sample use in the reference, LabelGen.py:63-101, WorldFeatLabelGen.py:459-567) and the `wave` module."""
import os
import struct
import wave
import zipfile

import numpy as np
import pytest

from idiaptts_b200 import corpus_io as cio
from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen


def _write_wav(path, x, fs=22050, width=2, channels=1):
    with wave.open(path, "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(width)
        w.setframerate(fs)
        w.writeframes(x.tobytes())


def test_packed_wav_reader_equals_wave_module(tmp_path):
    rng = np.random.default_rng(1)
    paths, waves = [], []
    for i, n in enumerate([1, 2, 777, 48000, 143325, 0, 5]):
        x = rng.integers(-32768, 32768, n).astype(np.int16)
        p = str(tmp_path / ("u%d.wav" % i))
        _write_wav(p, x)
        paths.append(p)
        waves.append(x)
    info = cio.probe_wavs(paths, threads=3)
    assert info["num_samples"].tolist() == [len(w) for w in waves]
    assert set(info["fs"].tolist()) == {22050} and set(info["bits"].tolist()) == {16} and set(info["channels"].tolist()) == {1}
    for threads in (1, 0):
        samples, off, fs = cio.read_wavs_i16(paths, threads=threads)
        assert fs == 22050 and off.tolist() == np.concatenate(([0], np.cumsum([len(w) for w in waves]))).tolist()
        for i, w in enumerate(waves):
            assert np.array_equal(samples.numpy()[off[i]:off[i + 1]], w)
    assert cio.read_wavs_i16([])[1].tolist() == [0]


def test_wav_headers_with_extra_chunks_and_extensible_format(tmp_path):
    x = np.arange(-50, 50, dtype=np.int16)
    # LIST chunk of odd length (padded) before fmt / data, WAVE_FORMAT_EXTENSIBLE with the PCM sub-format
    fmt = struct.pack("<HHIIHHHHI", 0xFFFE, 1, 16000, 32000, 2, 16, 22, 16, 4) + struct.pack("<H", 1) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    body = b"WAVE" + b"LIST" + struct.pack("<I", 5) + b"hello\x00" + b"fmt " + struct.pack("<I", len(fmt)) + fmt
    body += b"data" + struct.pack("<I", x.nbytes) + x.tobytes() + b"junk" + struct.pack("<I", 2) + b"zz"
    p = str(tmp_path / "ext.wav")
    with open(p, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)
    info = cio.probe_wavs([p])
    assert (info["num_samples"][0], info["fs"][0], info["bits"][0], info["channels"][0]) == (100, 16000, 16, 1)
    samples, _, fs = cio.read_wavs_i16([p])
    assert fs == 16000 and np.array_equal(samples.numpy(), x)
    with wave.open(p, "rb") as w:  # the Python reader sees the same file
        assert w.getnframes() == 100 and np.array_equal(np.frombuffer(w.readframes(100), np.int16), x)


def test_wav_reader_refuses_what_it_does_not_take(tmp_path):
    a, b, c = str(tmp_path / "a.wav"), str(tmp_path / "b.wav"), str(tmp_path / "c.wav")
    _write_wav(a, np.zeros(10, np.int16))
    _write_wav(b, np.zeros(10, np.int32), width=4)
    _write_wav(c, np.zeros(10, np.int16), fs=16000)
    info = cio.probe_wavs([a, b])
    assert info["bits"].tolist() == [16, 32]                 # the probe reports, the caller decides
    with pytest.raises(ValueError, match="32-bit"):
        cio.read_wavs_i16([a, b])
    with pytest.raises(ValueError, match="mixed sampling rates"):
        cio.read_wavs_i16([a, c])
    with pytest.raises(ValueError, match="cannot open"):
        cio.probe_wavs([a, str(tmp_path / "missing.wav")])
    notwav = str(tmp_path / "n.wav")
    open(notwav, "wb").write(b"x" * 100)
    with pytest.raises(ValueError, match="not a RIFF/WAVE"):
        cio.probe_wavs([notwav])
    # IEEE-float wav (format tag 3) is not PCM
    f = str(tmp_path / "f.wav")
    fmt = struct.pack("<HHIIHH", 3, 1, 16000, 64000, 4, 32)
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"data" + struct.pack("<I", 8) + b"\x00" * 8
    open(f, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    with pytest.raises(ValueError, match="format tag 3"):
        cio.probe_wavs([f])


@pytest.mark.parametrize("threads", [1, 0])
def test_npz_writer_is_read_by_numpy(tmp_path, threads):
    rng = np.random.default_rng(2)
    rows = [0, 1, 13, 1301, 400]
    off = np.concatenate(([0], np.cumsum(rows)))
    feats = rng.standard_normal((off[-1], 192)).astype(np.float32)
    feats[3, 5] = np.nan
    feats[4, 6] = -np.inf
    paths = [str(tmp_path / ("u%d.npz" % i)) for i in range(len(rows))]
    keys = ["mcep", "mcep_deltas", "mcep_double_deltas"]
    cio.write_npz(paths, keys, [0, 64, 128], [60, 60, 60], off, feats, threads=threads)
    for i, p in enumerate(paths):
        assert zipfile.ZipFile(p).testzip() is None          # CRC-32 of every member verified by Python's zipfile
        with np.load(p) as arc:
            assert arc.files == keys
            for k, c0 in zip(keys, (0, 64, 128)):
                a = arc[k]
                assert a.dtype == np.float32 and a.shape == (rows[i], 60) and a.flags["C_CONTIGUOUS"]
                assert np.array_equal(a, feats[off[i]:off[i + 1], c0:c0 + 60], equal_nan=True)
    # single-column and full-width (contiguous) blocks, torch tensors as the source
    import torch
    t = torch.from_numpy(feats)
    cio.write_npz(paths, ["vuv"], [61], [1], off, t, threads=threads)
    cio.write_npz([p + ".full.npz" for p in paths], ["all"], [0], [192], off, t, threads=threads)
    with np.load(paths[3]) as arc:
        assert arc.files == ["vuv"] and np.array_equal(arc["vuv"], feats[off[3]:off[4], 61:62])
    with np.load(paths[3] + ".full.npz") as arc:
        assert np.array_equal(arc["all"], feats[off[3]:off[4]], equal_nan=True)
    # byte-for-byte the member numpy.save would produce
    import io
    buf = io.BytesIO()
    np.save(buf, feats[off[3]:off[4], 61:62])
    assert zipfile.ZipFile(paths[3]).read("vuv.npy") == buf.getvalue()


def test_npz_reader_takes_numpy_archives(tmp_path):
    rng = np.random.default_rng(3)
    rows = [5, 0, 300, 1301]
    off = np.concatenate(([0], np.cumsum(rows)))
    feats = rng.standard_normal((off[-1], 10)).astype(np.float32)
    paths = []
    for i in range(len(rows)):
        r = feats[off[i]:off[i + 1]]
        np.savez(str(tmp_path / ("n%d" % i)), mcep=r[:, :6], mcep_deltas=r[:, 6:9], lf0=r[:, 9], other=np.arange(4))  # lf0 is 1-D here
        paths.append(str(tmp_path / ("n%d.npz" % i)))
    r6, c6 = cio.probe_npz(paths, "mcep")
    assert r6.tolist() == rows and c6.tolist() == [6] * 4
    r1, c1 = cio.probe_npz(paths, "lf0")
    assert r1.tolist() == rows and c1.tolist() == [1] * 4
    back = np.full((off[-1], 12), 7.0, np.float32)
    cio.read_npz(paths, ["mcep", "mcep_deltas", "lf0"], [0, 8, 11], [6, 3, 1], off, back, threads=2)
    assert np.array_equal(back[:, 0:6], feats[:, :6]) and np.array_equal(back[:, 8:11], feats[:, 6:9])
    assert np.array_equal(back[:, 11], feats[:, 9]) and np.all(back[:, 6:8] == 7.0)
    # shape mismatch, missing key, wrong dtype, compressed archive, corrupt data: refused, with the file named
    with pytest.raises(ValueError, match="expected"):
        cio.read_npz(paths, ["mcep"], [0], [5], off, back)
    with pytest.raises(ValueError, match="no array 'nope'"):
        cio.probe_npz(paths, "nope")
    with pytest.raises(ValueError, match="float32"):
        cio.probe_npz(paths, "other")
    np.savez_compressed(str(tmp_path / "c"), mcep=feats[:5, :6])
    with pytest.raises(ValueError, match="compressed"):
        cio.probe_npz([str(tmp_path / "c.npz")], "mcep")
    raw = bytearray(open(paths[3], "rb").read())
    raw[len(raw) // 2] ^= 0x40
    open(paths[3], "wb").write(raw)
    with pytest.raises(ValueError, match="CRC"):
        cio.read_npz(paths, ["mcep", "mcep_deltas", "lf0"], [0, 8, 11], [6, 3, 1], off, back)
    cio.read_npz(paths, ["mcep"], [0], [6], off, back, verify_crc=False)   # numpy.load(..., mmap) semantics: no check


def test_round_trip_and_load_batch_equal_the_per_sample_reader(tmp_path):
    """write_npz -> WorldFeatLabelGen.load (numpy) and -> load_batch (native) give the same rows; load_batch falls back to the
    per-sample reader for archives the native reader refuses."""
    rng = np.random.default_rng(4)
    D, nap = 6, 2
    for add_deltas in (False, True):
        root = str(tmp_path / ("lab%d" % add_deltas))
        ids = ["a", "bb", "ccc", "dddd"]
        rows = [30, 1, 77, 5]
        off = np.concatenate(([0], np.cumsum(rows)))
        dim = D + 2 + nap
        W = dim * (3 if add_deltas else 1)
        feats = rng.standard_normal((off[-1], W)).astype(np.float32)
        groups = (("mcep%d" % D, "mcep", 0, D, True), ("lf0", "lf0", D, 1, True), ("vuv", "vuv", D + 1, 1, False), ("bap", "bap", D + 2, nap, True))
        expect = []
        for sub, key, c0, d, deltas in groups:
            os.makedirs(os.path.join(root, sub))
            nblk = 3 if (add_deltas and deltas) else 1
            keys = [key, key + "_deltas", key + "_double_deltas"][:nblk]
            cio.write_npz([os.path.join(root, sub, i + ".npz") for i in ids], keys, [c0 + b * dim for b in range(nblk)], [d] * nblk, off, feats)
            expect += [np.arange(c0 + b * dim, c0 + b * dim + d) for b in range(nblk)]
        expect = np.concatenate(expect)
        reader = WorldFeatLabelGen(root, add_deltas=add_deltas, num_coded_sps=D, num_bap=nap, io_threads=2)
        for u, i in enumerate(ids):
            assert np.array_equal(reader.load(i), feats[off[u]:off[u + 1]][:, expect])
        got, goff = reader.load_batch(ids, pin=False)
        assert goff.tolist() == off.tolist() and np.array_equal(got.numpy(), feats[:, expect])
        got2, goff2 = reader.load_batch(["bb.npz", "/x/y/a"], pin=False)       # ids are reduced to base names without extension
        assert goff2.tolist() == [0, 1, 31] and np.array_equal(got2.numpy(), np.concatenate((feats[30:31], feats[0:30]))[:, expect])
        # one compressed archive in the set: the whole call goes through the numpy reader, same result
        with np.load(os.path.join(root, "lf0", "bb.npz")) as arc:
            content = {k: arc[k] for k in arc.files}
        np.savez_compressed(os.path.join(root, "lf0", "bb"), **content)
        got3, goff3 = reader.load_batch(ids, pin=False)
        assert goff3.tolist() == off.tolist() and np.array_equal(got3.numpy(), feats[:, expect])
        with pytest.raises(FileNotFoundError):
            reader.load_batch(["a", "missing"], pin=False)


def test_crc32_variants_equal_zlib():
    """The archives' CRC-32: the table code and the carry-less-multiplication folding (taken where the CPU has PCLMULQDQ) against
    zlib, for every length around the 16 / 64-byte block boundaries and unaligned starts."""
    import zlib
    from idiaptts_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(9)
    for n in list(range(0, 200)) + [255, 256, 257, 1023, 4096, 65537, 1000003]:
        b = rng.integers(0, 256, n + 3, dtype=np.uint8)
        for off in (0, 1, 3):
            v = b[off:off + n]
            want = zlib.crc32(v.tobytes())
            assert lib.b2w_crc32(v.ctypes.data, n, 0) == want and lib.b2w_crc32(v.ctypes.data, n, 1) == want, (n, off)
