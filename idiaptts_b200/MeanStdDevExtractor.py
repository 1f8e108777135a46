"""Online mean / standard deviation of feature streams, file formats and combination rules of the reference's
idiaptts/misc/normalisation/MeanStdDevExtractor.py (add_sample :43-47, get_params :49-53, save :55-98, load :117-160,
combine_stats :163-204, combine_mean_std :207-255), plus the entry point the GPU path uses: the per-column sums arrive
already reduced (fp64, one NCCL all-reduce across ranks) through `add_sums`.

Differences kept deliberately small: the reference accumulates in the sample dtype (float32 for WORLD features); sums
produced on the GPU are fp64 (agreement ~3e-6).  `np.str` / `np.int` (removed from numpy) are spelled
`str` / `int`."""
import logging
import os
import struct

import numpy as np


class MeanStdDevExtractor(object):
    logger = logging.getLogger(__name__)

    file_name_stats = "stats"
    file_name_appendix = "mean-std_dev"

    def __init__(self):
        self.sum_length = 0
        self.sum_frames = 0
        self.sum_squared_frames = 0

    def _normalise(self, feature, mean, std_dev):
        return (feature - mean) / std_dev

    def _denormalise(self, feature, mean, std_dev):
        return feature * std_dev + mean

    def add_sample(self, sample):
        assert sample is not None, "Sample cannot be None."
        self.sum_length += len(sample)
        self.sum_frames += np.sum(sample, axis=0)
        self.sum_squared_frames += np.sum(sample ** 2, axis=0)

    def add_sums(self, length, sum_frames, sum_squared_frames):
        """Pre-reduced statistics of `length` frames (what the CUDA statistics kernel + all-reduce deliver)."""
        self.sum_length += int(length)
        self.sum_frames = self.sum_frames + np.asarray(sum_frames, np.float64)
        self.sum_squared_frames = self.sum_squared_frames + np.asarray(sum_squared_frames, np.float64)

    def get_params(self):
        mean = self.sum_frames / self.sum_length
        std_dev = np.sqrt(self.sum_squared_frames / self.sum_length - mean ** 2)
        return np.atleast_1d(mean), np.atleast_1d(std_dev)

    def save(self, filename, datatype=np.float64):
        self.save_stats(filename, datatype)
        self.save_mean_std_dev(filename, datatype)

    def save_stats(self, filename, datatype=np.float64):
        if filename is not None and os.path.basename(filename) != "":
            filename += "-"
        self._save(filename + self.file_name_stats, self.sum_length,
                   {"sum_frames": self.sum_frames, "sum_squared_frames": self.sum_squared_frames}, datatype)

    def save_mean_std_dev(self, filename, datatype=np.float64):
        if filename is not None and os.path.basename(filename) != "":
            filename += "-"
        mean, std_dev = self.get_params()
        self._save(filename + self.file_name_appendix, self.sum_length, {"mean": mean, "std_dev": std_dev}, datatype)

    @staticmethod
    def _save(filename, sum_length, stats, datatype):
        if datatype is str:
            np.savetxt(filename + ".txt", np.concatenate([np.atleast_2d(v) for v in stats.values()], axis=0),
                       header=str(sum_length))
        elif datatype is np.float32 or datatype is np.float64:
            out = {k: np.atleast_1d(v).astype(datatype, copy=False) for k, v in stats.items()}
            out["sum_length"] = np.array(sum_length, dtype=int)
            np.savez(filename, **out)
        else:
            logging.error("Unknown datatype: {}. Please choose one of [numpy.float32, numpy.float64, str].".format(datatype))

    @staticmethod
    def load_stats(file_path, datatype=np.float64):
        if datatype is str:
            with open(file_path, "r") as f:
                labels_len = int(f.readline().lstrip("# "))
                stats = np.loadtxt(f, dtype=np.float64)
            sum_frames, sum_squared_frames = np.split(stats, stats.shape[0], axis=0)
        else:
            archive = np.load(file_path)
            labels_len = archive["sum_length"]
            sum_frames = archive["sum_frames"]
            sum_squared_frames = archive["sum_squared_frames"]
        return sum_frames, sum_squared_frames, labels_len

    @staticmethod
    def load(file_path, datatype=np.float64):
        if datatype is str:
            with open(file_path, "r") as f:
                f.readline()
                mean_std_dev = np.loadtxt(f, dtype=np.float32)
            mean, std_dev = np.split(mean_std_dev, mean_std_dev.shape[0], axis=0)
        elif file_path.endswith(".bin"):  # legacy: int32 N + float64 [2 x d]
            with open(file_path, "rb") as f:
                struct.unpack("i", f.read(4))
                mean_std_dev = np.fromfile(f, dtype=datatype).reshape((2, -1))
            mean, std_dev = np.split(mean_std_dev, mean_std_dev.shape[0], axis=0)
        else:
            archive = np.load(file_path)
            mean, std_dev = archive["mean"], archive["std_dev"]
        return mean.astype(np.float32, copy=False), std_dev.astype(np.float32, copy=False)

    @staticmethod
    def load_mean_std_dev_from_stats(file_path, datatype=np.float64):
        sum_frames, sum_squared_frames, sum_length = MeanStdDevExtractor.load_stats(file_path, datatype)
        mean = sum_frames / sum_length
        std_dev = np.sqrt(sum_squared_frames / sum_length - mean ** 2)
        return mean.astype(np.float32, copy=False), std_dev.astype(np.float32, copy=False)

    @staticmethod
    def combine_stats(file_list, dir_out=None, datatype=np.float64, save_txt=False):
        """Sum of the per-subset sums: exactly what the NCCL all-reduce of the statistics buffer computes across ranks."""
        sum_length = 0
        sum_frames = 0
        sum_squared_frames = 0
        for file in file_list:
            cur_sum, cur_sq, labels_len = MeanStdDevExtractor.load_stats(file, datatype=datatype)
            sum_length += labels_len
            sum_frames += cur_sum
            sum_squared_frames += cur_sq
        if dir_out is not None:
            filename = os.path.join(dir_out, MeanStdDevExtractor.file_name_stats)
            stats = {"sum_frames": sum_frames, "sum_squared_frames": sum_squared_frames}
            MeanStdDevExtractor._save(filename, sum_length, stats, datatype=np.float32)
            if save_txt:
                MeanStdDevExtractor._save(filename, sum_length, stats, datatype=str)
        return sum_length, sum_frames, sum_squared_frames

    @staticmethod
    def combine_mean_std(file_list, dir_out=None, datatype=np.float64, save_txt=True):
        sum_length, sum_frames, sum_squared_frames = MeanStdDevExtractor.combine_stats(file_list, dir_out=dir_out,
                                                                                       datatype=datatype)
        mean = np.atleast_2d(sum_frames / sum_length)
        variance = np.atleast_2d(sum_squared_frames / sum_length) - mean ** 2
        negative = (variance < 0)[0]
        if negative.any():
            logging.warning("Encountered negative variance for indices {} when combining statistics of {}. Setting those "
                            "elements to 0 instead.".format(np.arange(variance.shape[1])[negative], file_list))
            variance[:, negative] = 0.0
        std_dev = np.sqrt(variance)
        if dir_out is not None:
            filename = os.path.join(dir_out, MeanStdDevExtractor.file_name_appendix)
            stats = {"mean": mean, "std_dev": std_dev}
            MeanStdDevExtractor._save(filename, sum_length, stats, datatype=datatype)
            if save_txt:
                MeanStdDevExtractor._save(filename, sum_length, stats, datatype=str)
        return mean, std_dev
