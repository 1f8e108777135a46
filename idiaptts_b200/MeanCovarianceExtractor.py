"""Online mean / covariance of feature streams (the add_deltas recipe): sums and Gram matrix, as the reference's
idiaptts/misc/normalisation/MeanCovarianceExtractor.py (add_sample :44-48, get_params :50-55, combine_stats :168-217).
The GPU path delivers sum x and X^T X already reduced (fp64) through `add_sums`."""
import os

import numpy as np


class MeanCovarianceExtractor(object):
    file_name_stats = "stats"
    file_name_appendix = "mean-covariance"

    def __init__(self):
        self.sum_length = 0
        self.sum_frames = None
        self.sum_product_frames = None

    def add_sample(self, sample):
        sample = np.asarray(sample)
        if sample.ndim == 1:
            sample = sample[:, None]
        self.add_sums(len(sample), np.sum(sample, axis=0, keepdims=True), np.dot(sample.T, sample))

    def add_sums(self, length, sum_frames, sum_product_frames):
        sum_frames = np.atleast_2d(np.asarray(sum_frames, np.float64))
        sum_product_frames = np.asarray(sum_product_frames, np.float64)
        self.sum_length += int(length)
        self.sum_frames = sum_frames if self.sum_frames is None else self.sum_frames + sum_frames
        self.sum_product_frames = (sum_product_frames if self.sum_product_frames is None
                                   else self.sum_product_frames + sum_product_frames)

    def get_params(self):
        mean = self.sum_frames / self.sum_length
        covariance = self.sum_product_frames / self.sum_length - np.dot(mean.T, mean)
        std_dev = np.sqrt(np.diag(covariance))[None, :]
        return mean, covariance, std_dev

    def save(self, filename, datatype=np.float64):
        if filename is not None and os.path.basename(filename) != "":
            filename += "-"
        np.savez(filename + self.file_name_stats, sum_frames=self.sum_frames.astype(datatype),
                 sum_product_frames=self.sum_product_frames.astype(datatype), sum_length=np.array(self.sum_length, dtype=int))
        mean, covariance, std_dev = self.get_params()
        np.savez(filename + self.file_name_appendix, mean=mean.astype(datatype), covariance=covariance.astype(datatype),
                 std_dev=std_dev.astype(datatype), sum_length=np.array(self.sum_length, dtype=int))

    @staticmethod
    def load(file_path, datatype=np.float64):
        archive = np.load(file_path)
        mean, covariance = archive["mean"], archive["covariance"]
        std_dev = archive["std_dev"] if "std_dev" in archive else np.sqrt(np.diag(covariance))[None, :]
        return mean.astype(np.float32), covariance.astype(np.float32), std_dev.astype(np.float32)

    @staticmethod
    def combine_stats(file_list, dir_out=None, datatype=np.float64):
        total = MeanCovarianceExtractor()
        for file in file_list:
            archive = np.load(file)
            total.add_sums(int(archive["sum_length"]), archive["sum_frames"], archive["sum_product_frames"])
        if dir_out is not None:
            total.save(os.path.join(dir_out, ""), datatype)
        return total.sum_length, total.sum_frames, total.sum_product_frames

    @staticmethod
    def combine_mean_covariance(file_list, dir_out=None, datatype=np.float64):
        total = MeanCovarianceExtractor()
        for file in file_list:
            archive = np.load(file)
            total.add_sums(int(archive["sum_length"]), archive["sum_frames"], archive["sum_product_frames"])
        if dir_out is not None:
            total.save(os.path.join(dir_out, ""), datatype)
        return total.get_params()
