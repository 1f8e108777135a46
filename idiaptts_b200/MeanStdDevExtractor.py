"""Corpus mean / standard deviation for feature normalisation: the accumulator, its on-disk formats and the rules for merging
subsets that the reference defines in idiaptts/misc/normalisation/MeanStdDevExtractor.py.

The class is organised around what crosses the boundary rather than around the reference's text:

  state      N (frames), sum x [d], sum x^2 [d], held in float64.  The CUDA statistics kernel and the NCCL all-reduce deliver
             exactly these three (`add_sums`); `add_sample` reduces a host array to them.  (The reference accumulates in the
             sample dtype, float32 for WORLD features: agreement ~3e-6)
  parameters mean = sum x / N, std_dev = sqrt(sum x^2 / N - mean^2)                                   (reference :49-53)
  files      <prefix->stats.npz         {sum_frames, sum_squared_frames, sum_length}                  (reference :59-98)
             <prefix->mean-std_dev.npz  {mean, std_dev, sum_length}
             .txt variants (header = N, one row per array) when datatype is str; legacy .bin = int32 N + raw [2, d] floats
  merging    sums of subsets add; a negative variance from cancellation is clamped to zero           (reference :163-255)
"""
import logging
import os
import struct

import numpy as np

_FLOAT_TYPES = (np.float32, np.float64)


def _prefix(filename):
    """'<dir>/<name>' -> '<dir>/<name>-', '<dir>/' -> '<dir>/' (the reference's file naming rule)."""
    filename = "" if filename is None else str(filename)
    return filename + "-" if os.path.basename(filename) != "" else filename


def _write_arrays(path, sum_length, arrays, datatype):
    """One statistics file: npz (float32 / float64) or text (datatype str)."""
    if datatype is str:
        np.savetxt(path + ".txt", np.concatenate([np.atleast_2d(a) for a in arrays.values()], axis=0), header=str(sum_length))
    elif datatype in _FLOAT_TYPES:
        payload = {k: np.atleast_1d(v).astype(datatype, copy=False) for k, v in arrays.items()}
        payload["sum_length"] = np.array(sum_length, dtype=int)
        np.savez(path, **payload)
    else:
        logging.error("Unknown datatype: %s. Please choose one of [numpy.float32, numpy.float64, str].", getattr(datatype, "__name__", datatype))


def _read_text(path):
    with open(path, "r") as f:
        header = f.readline().lstrip("# ").strip()
        rows = np.loadtxt(f, dtype=np.float64, ndmin=2)
    return int(float(header)), rows


class MeanStdDevExtractor(object):
    logger = logging.getLogger(__name__)

    file_name_stats = "stats"
    file_name_appendix = "mean-std_dev"
    _second_moment_key = "sum_squared_frames"

    def __init__(self):
        self.sum_length = 0
        self.sum_frames = 0
        self.sum_squared_frames = 0

    # ---- normalisation itself (used by the readers) ------------------------------------------------------------------------
    def _normalise(self, feature, mean, std_dev):
        return (feature - mean) / std_dev

    def _denormalise(self, feature, mean, std_dev):
        return feature * std_dev + mean

    # ---- accumulation ---------------------------------------------------------------------------------------------------------
    def add_sums(self, length, sum_frames, sum_squared_frames):
        """Pre-reduced statistics of `length` frames (what the CUDA statistics kernel + all-reduce deliver)."""
        self.sum_length += int(length)
        self.sum_frames = self.sum_frames + np.asarray(sum_frames, np.float64)
        self.sum_squared_frames = self.sum_squared_frames + np.asarray(sum_squared_frames, np.float64)

    def add_sample(self, sample):
        assert sample is not None, "Sample cannot be None."
        x = np.asarray(sample, np.float64)
        self.add_sums(len(x), x.sum(axis=0), np.square(x).sum(axis=0))

    @staticmethod
    def _params_from_sums(sum_length, sum_frames, sum_squared_frames, clamp=False):
        mean = sum_frames / sum_length
        variance = sum_squared_frames / sum_length - mean ** 2
        if clamp and np.any(variance < 0):
            bad = np.nonzero(np.atleast_1d(variance < 0).reshape(-1))[0]
            logging.warning("Encountered negative variance for indices %s; setting those elements to 0 instead.", bad)
            variance = np.where(variance < 0, 0.0, variance)
        return mean, np.sqrt(variance)

    def get_params(self):
        mean, std_dev = self._params_from_sums(self.sum_length, self.sum_frames, self.sum_squared_frames)
        return np.atleast_1d(mean), np.atleast_1d(std_dev)

    # ---- files ------------------------------------------------------------------------------------------------------------------
    def save(self, filename, datatype=np.float64):
        self.save_stats(filename, datatype)
        self.save_mean_std_dev(filename, datatype)

    def save_stats(self, filename, datatype=np.float64):
        self._save(_prefix(filename) + self.file_name_stats, self.sum_length,
                   {"sum_frames": self.sum_frames, "sum_squared_frames": self.sum_squared_frames}, datatype)

    def save_mean_std_dev(self, filename, datatype=np.float64):
        mean, std_dev = self.get_params()
        self._save(_prefix(filename) + self.file_name_appendix, self.sum_length, {"mean": mean, "std_dev": std_dev}, datatype)

    @staticmethod
    def _save(filename, sum_length, stats, datatype):
        _write_arrays(filename, sum_length, stats, datatype)

    @staticmethod
    def load_stats(file_path, datatype=np.float64):
        """-> (sum_frames, sum_squared_frames, sum_length)"""
        if datatype is str:
            n, rows = _read_text(file_path)
            return rows[0:1], rows[1:2], n
        if datatype not in _FLOAT_TYPES:
            logging.error("Unknown datatype: %s.", getattr(datatype, "__name__", datatype))
            return None
        with np.load(file_path) as arc:
            return arc["sum_frames"], arc["sum_squared_frames"], arc["sum_length"]

    @staticmethod
    def load(file_path, datatype=np.float64):
        """-> (mean, std_dev) as float32."""
        if datatype is str:
            _, rows = _read_text(file_path)
            mean, std_dev = rows[0:1], rows[1:2]
        elif datatype not in _FLOAT_TYPES:
            logging.error("Unknown datatype: %s.", getattr(datatype, "__name__", datatype))
            return None
        elif str(file_path).endswith(".bin"):  # legacy: int32 frame count, then [2, d] raw values
            with open(file_path, "rb") as f:
                struct.unpack("i", f.read(4))
                both = np.fromfile(f, dtype=datatype).reshape((2, -1))
            mean, std_dev = both[0:1], both[1:2]
        else:
            with np.load(file_path) as arc:
                mean, std_dev = arc["mean"], arc["std_dev"]
        return mean.astype(np.float32, copy=False), std_dev.astype(np.float32, copy=False)

    @staticmethod
    def load_mean_std_dev_from_stats(file_path, datatype=np.float64):
        sum_frames, sum_squared_frames, sum_length = MeanStdDevExtractor.load_stats(file_path, datatype)
        mean, std_dev = MeanStdDevExtractor._params_from_sums(sum_length, sum_frames, sum_squared_frames)
        return mean.astype(np.float32, copy=False), std_dev.astype(np.float32, copy=False)

    # ---- merging subsets (the reference's only "collective": file based; the GPU path does the same sum as one all-reduce) ----
    @staticmethod
    def combine_stats(file_list, dir_out=None, datatype=np.float64, save_txt=False):
        total = MeanStdDevExtractor()
        for file in file_list:
            s, q, n = MeanStdDevExtractor.load_stats(file, datatype=datatype)
            total.add_sums(int(n), s, q)
        if dir_out is not None:
            path = os.path.join(dir_out, MeanStdDevExtractor.file_name_stats)
            stats = {"sum_frames": total.sum_frames, "sum_squared_frames": total.sum_squared_frames}
            _write_arrays(path, total.sum_length, stats, np.float32)
            if save_txt:
                _write_arrays(path, total.sum_length, stats, str)
        return total.sum_length, total.sum_frames, total.sum_squared_frames

    @staticmethod
    def combine_mean_std(file_list, dir_out=None, datatype=np.float64, save_txt=True):
        sum_length, sum_frames, sum_squared_frames = MeanStdDevExtractor.combine_stats(file_list, dir_out=dir_out, datatype=datatype)
        mean, std_dev = MeanStdDevExtractor._params_from_sums(sum_length, sum_frames, sum_squared_frames, clamp=True)
        if dir_out is not None:
            path = os.path.join(dir_out, MeanStdDevExtractor.file_name_appendix)
            _write_arrays(path, sum_length, {"mean": mean, "std_dev": std_dev}, datatype)
            if save_txt:
                _write_arrays(path, sum_length, {"mean": mean, "std_dev": std_dev}, str)
        return mean, std_dev
