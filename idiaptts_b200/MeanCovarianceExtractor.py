"""Corpus mean / covariance (the add_deltas recipe: MLPG needs the covariance of [static | delta | delta-delta]): accumulator,
file formats and merging rules of the reference's idiaptts/misc/normalisation/MeanCovarianceExtractor.py, organised like
MeanStdDevExtractor.py in this package:

  state      N, sum x [1, d], X^T X [d, d] in float64 (`add_sums`: the CUDA Gram kernel + all-reduce deliver these)
  parameters mean = sum x / N, covariance = X^T X / N - mean^T mean; get_params() -> [mean, covariance] (reference :50-55)
  files      <prefix->stats.npz            {sum_frames, sum_product_frames, sum_length}               (reference :61-98)
             <prefix->mean-covariance.npz  {mean, covariance, sum_length}
             legacy .bin = int32 N, int32 rows, then raw [rows, d] values: first row the mean, the rest the covariance
  load()     -> (mean [d], covariance [d, d], std_dev [d]) as float32, std_dev = sqrt(diag(covariance)) (reference :120-150)
"""
import logging
import os
import struct

import numpy as np

from .MeanStdDevExtractor import _FLOAT_TYPES, _prefix, _read_text, _write_arrays


class MeanCovarianceExtractor(object):
    logger = logging.getLogger(__name__)

    file_name_stats = "stats"
    file_name_appendix = "mean-covariance"

    def __init__(self):
        self.sum_length = 0
        self.sum_frames = 0
        self.sum_product_frames = 0

    def _normalise(self, feature, mean, std_dev):
        return (feature - mean) / std_dev

    def _denormalise(self, feature, mean, std_dev):
        return feature * std_dev + mean

    def add_sums(self, length, sum_frames, sum_product_frames):
        self.sum_length += int(length)
        self.sum_frames = self.sum_frames + np.atleast_2d(np.asarray(sum_frames, np.float64))
        self.sum_product_frames = self.sum_product_frames + np.asarray(sum_product_frames, np.float64)

    def add_sample(self, sample):
        assert sample is not None, "Sample cannot be None."
        x = np.asarray(sample, np.float64)
        if x.ndim == 1:
            x = x[:, None]
        self.add_sums(len(x), x.sum(axis=0, keepdims=True), x.T @ x)

    @staticmethod
    def _params_from_sums(sum_length, sum_frames, sum_product_frames):
        mean = np.atleast_2d(sum_frames) / sum_length
        return mean, sum_product_frames / sum_length - mean.T @ mean

    def get_params(self):
        """[mean [1, d], covariance [d, d]] -- two values, like the reference."""
        mean, covariance = self._params_from_sums(self.sum_length, self.sum_frames, self.sum_product_frames)
        return np.atleast_2d(mean, covariance)

    def get_std_dev(self):
        return np.sqrt(np.diag(self.get_params()[1]))

    # ---- files ------------------------------------------------------------------------------------------------------------------
    def save(self, filename, datatype=np.float64):
        self.save_stats(filename, datatype)
        self.save_mean_covariance(filename, datatype)

    def save_stats(self, filename, datatype=np.float64):
        self._save(_prefix(filename) + self.file_name_stats, self.sum_length,
                   {"sum_frames": self.sum_frames, "sum_product_frames": self.sum_product_frames}, datatype)

    def save_mean_covariance(self, filename, datatype=np.float64):
        mean, covariance = self.get_params()
        self._save(_prefix(filename) + self.file_name_appendix, self.sum_length, {"mean": mean, "covariance": covariance}, datatype)

    @staticmethod
    def _save(filename, sum_length, stats, datatype):
        _write_arrays(filename, sum_length, stats, datatype)

    @staticmethod
    def load_stats(file_path, datatype=np.float64):
        """-> (sum_frames, sum_product_frames, sum_length)"""
        if datatype is str:
            n, rows = _read_text(file_path)
            return rows[0:1], rows[1:], n
        if datatype not in _FLOAT_TYPES:
            logging.error("Unknown datatype: %s.", getattr(datatype, "__name__", datatype))
            return None
        with np.load(file_path) as arc:
            return arc["sum_frames"], arc["sum_product_frames"], arc["sum_length"]

    @staticmethod
    def load(file_path, datatype=np.float64):
        """-> (mean [d], covariance [d, d], std_dev [d]) as float32."""
        if datatype is str:
            _, rows = _read_text(file_path)
            mean, covariance = rows[0:1], rows[1:]
        elif datatype not in _FLOAT_TYPES:
            logging.error("Unknown datatype: %s.", getattr(datatype, "__name__", datatype))
            return None
        elif str(file_path).endswith(".bin"):  # legacy: int32 frame count, int32 row count, then [rows, d] raw values
            with open(file_path, "rb") as f:
                _, rows = struct.unpack("ii", f.read(8))
                both = np.fromfile(f, dtype=datatype).reshape((rows, -1))
            mean, covariance = both[0:1], both[1:]
        else:
            with np.load(file_path) as arc:
                mean, covariance = arc["mean"], arc["covariance"]
        std_dev = np.sqrt(np.diag(np.atleast_2d(covariance))).astype(np.float32)
        return (np.squeeze(mean).astype(np.float32, copy=False), np.atleast_2d(covariance.astype(np.float32, copy=False)),
                np.squeeze(std_dev))

    @staticmethod
    def load_mean_covariance_from_stats(file_path, datatype=np.float64):
        sum_frames, sum_product_frames, sum_length = MeanCovarianceExtractor.load_stats(file_path, datatype)
        mean, covariance = MeanCovarianceExtractor._params_from_sums(sum_length, sum_frames, sum_product_frames)
        return np.atleast_2d(mean.astype(np.float32, copy=False)), covariance.astype(np.float32, copy=False)

    # ---- merging subsets -------------------------------------------------------------------------------------------------------
    @staticmethod
    def combine_stats(file_list, dir_out=None, file_name=None, datatype=np.float64, save_txt=False):
        total = MeanCovarianceExtractor()
        for file in file_list:
            s, g, n = MeanCovarianceExtractor.load_stats(file, datatype=datatype)
            total.add_sums(int(n), s, g)
        if dir_out is not None:
            path = os.path.join(dir_out, _prefix(file_name) + MeanCovarianceExtractor.file_name_stats)
            stats = {"sum_frames": total.sum_frames, "sum_product_frames": total.sum_product_frames}
            _write_arrays(path, total.sum_length, stats, datatype)
            if save_txt:
                _write_arrays(path, total.sum_length, stats, str)
        return total.sum_length, total.sum_frames, total.sum_product_frames

    @staticmethod
    def combine_mean_covariance(file_list, dir_out=None, file_name=None, datatype=np.float64, save_txt=True):
        sum_length, sum_frames, sum_product_frames = MeanCovarianceExtractor.combine_stats(file_list, dir_out=dir_out,
                                                                                           file_name=file_name, datatype=datatype)
        mean, covariance = MeanCovarianceExtractor._params_from_sums(sum_length, sum_frames, sum_product_frames)
        if dir_out is not None:
            path = os.path.join(dir_out, _prefix(file_name) + MeanCovarianceExtractor.file_name_appendix)
            _write_arrays(path, sum_length, {"mean": mean, "covariance": covariance}, datatype)
            if save_txt:
                _write_arrays(path, sum_length, {"mean": mean, "covariance": covariance}, str)
        return mean, covariance
