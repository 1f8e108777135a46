"""LF0LabelGen with the reference's entry points (idiaptts/src/data_preparation/world/LF0LabelGen.py), the F0 stage computed
by libb200world.so on the GPU:

    gen_data :212-321 (pyworld.dio + pyworld.stonemask per file, :263-264 -> ONE ragged GPU batch here), load_sample :121-136,
    load_lf0 :138-149, load_vuv :151-158, convert_to_world_features :160-167, get_normalisation_params :169-210,
    preprocess_sample / postprocess_sample :76-118, trim_end_sample :63-74.

On-disk formats as in the reference: raw float32 `<dir_out>/lf0/<id>.lf0` [T], `<dir_out>/vuv/<id>.vuv` [T], with deltas
`<dir_out>/lf0/<id>.lf0_deltas` [T x 4] = [lf0, d, dd, vuv]; statistics through MeanStdDevExtractor."""
import glob
import logging
import os
from collections import OrderedDict

import numpy as np
import torch

from . import ops
from .AudioProcessing import AudioProcessing
from .MeanStdDevExtractor import MeanStdDevExtractor


class LF0LabelGen:
    """Create LF0 feature labels for .wav files."""
    f0_silence_threshold = 20
    lf0_zero = 0

    dir_lf0 = "lf0"
    dir_deltas = "lf0"
    dir_vuv = "vuv"
    ext_lf0 = ".lf0"
    ext_deltas = ".lf0_deltas"
    ext_vuv = ".vuv"

    logger = logging.getLogger(__name__)

    def __init__(self, dir_labels, add_deltas=False):
        self.dir_labels = dir_labels
        self.add_deltas = add_deltas
        self.norm_params = None

    def __getitem__(self, id_name):
        return self.preprocess_sample(self.load_sample(id_name, self.dir_labels))

    @staticmethod
    def trim_end_sample(sample, length, reverse=False):
        if length == 0:
            return sample
        return sample[length:, ...] if reverse else sample[:-length, ...]

    def _params(self, norm_params):
        if norm_params is not None:
            return norm_params
        if self.norm_params is not None:
            return self.norm_params
        self.logger.error("Please give norm_params argument or call get_normaliations_params() before.")
        return None

    def preprocess_sample(self, sample, norm_params=None):
        p = self._params(norm_params)
        if p is None:
            return None
        return np.float32((sample - p[0]) / p[1])

    def postprocess_sample(self, sample, norm_params=None):
        p = self._params(norm_params)
        if p is None:
            return None
        return np.copy((sample * p[1]) + p[0])

    @staticmethod
    def load_sample(id_name, dir_out, add_deltas=False):
        lf0 = LF0LabelGen.load_lf0(id_name, dir_out, add_deltas)
        vuv = LF0LabelGen.load_vuv(id_name, dir_out)
        return np.concatenate((lf0, vuv), axis=1)

    @staticmethod
    def load_lf0(id_name, dir_out, add_deltas=False):
        ext, width = (LF0LabelGen.ext_deltas, 3) if add_deltas else (LF0LabelGen.ext_lf0, 1)
        with open(os.path.join(dir_out, LF0LabelGen.dir_lf0, id_name + ext), "rb") as f:
            return np.reshape(np.fromfile(f, dtype=np.float32), [-1, width])

    @staticmethod
    def load_vuv(id_name, dir_out):
        with open(os.path.join(dir_out, LF0LabelGen.dir_vuv, id_name + LF0LabelGen.ext_vuv), "rb") as f:
            return np.reshape(np.fromfile(f, dtype=np.float32), [-1, 1])

    @staticmethod
    def convert_to_world_features(sample):
        lf0 = sample[:, 0]
        vuv = np.copy(sample[:, -1])
        vuv[vuv < 0.5] = 0.0
        vuv[vuv >= 0.5] = 1.0
        return lf0, vuv

    def get_normalisation_params(self, dir_out, file_name=None):
        """Reads <file_name->mean-std_dev.npz (or the reference's legacy .bin) and stores (mean, std_dev) in self.norm_params."""
        prefix = (file_name + "-" if file_name is not None else "") + MeanStdDevExtractor.file_name_appendix
        sub = self.dir_deltas if self.add_deltas else self.dir_lf0
        path = os.path.join(dir_out, sub, prefix + ".npz")
        if not os.path.exists(path):
            path = os.path.join(dir_out, sub, prefix + ".bin")
        mean, std_dev = MeanStdDevExtractor.load(path)
        if not self.add_deltas:  # vuv: mean 0, std 1 (not saved by gen_data)
            mean = np.concatenate((np.atleast_2d(mean), np.atleast_2d(0.0)), axis=1)
            std_dev = np.concatenate((np.atleast_2d(std_dev), np.atleast_2d(1.0)), axis=1)
        self.norm_params = mean, std_dev
        return self.norm_params

    def gen_data(self, dir_in, dir_out=None, file_id_list="", id_list=None, add_deltas=False, return_dict=False):
        """LF0 and V/UV labels of all utterances in id_list, extracted as ONE ragged GPU batch (DIO + StoneMask + lf0 / vuv
        preparation + deltas); returns ([label_dict,] mean, std_dev) as the reference does."""
        if id_list is None:
            id_list = [os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(dir_in, "*.wav"))]
            file_id_list_name = "all"
        else:
            file_id_list_name = os.path.splitext(os.path.basename(file_id_list))[0]
        if dir_out is not None:
            if add_deltas:
                os.makedirs(os.path.join(dir_out, LF0LabelGen.dir_deltas), exist_ok=True)
            else:
                os.makedirs(os.path.join(dir_out, LF0LabelGen.dir_lf0), exist_ok=True)
                os.makedirs(os.path.join(dir_out, LF0LabelGen.dir_vuv), exist_ok=True)
        if not torch.cuda.is_available():
            raise RuntimeError("idiaptts_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        label_dict = OrderedDict()
        ext_lf0, ext_deltas = MeanStdDevExtractor(), MeanStdDevExtractor()
        # one batch per sampling rate (the reference accepts mixed rates: it works file by file)
        by_fs = OrderedDict()
        for name in id_list:
            x, fs = AudioProcessing.read_wav(os.path.join(dir_in, name + ".wav"))
            if x.dtype != np.int16:
                x = x.astype(np.float64)
            by_fs.setdefault((fs, x.dtype), []).append((name, x))
        results = {}
        for (fs, _), items in by_fs.items():
            waves = [x for _, x in items]
            f0s = [np.zeros(ops.num_frames(len(x), fs)) for x in waves]
            batch = ops.RaggedBatch.from_host(waves, f0s, fs, device=dev)
            ops.estimate_f0(batch)
            lf0, vuv = ops.lf0_vuv(batch.f0, batch.frame_off, LF0LabelGen.f0_silence_threshold, LF0LabelGen.lf0_zero)
            if add_deltas:
                d, dd = ops.deltas(lf0, batch.frame_off)
                rows = torch.cat((lf0, d, dd, vuv), dim=1).cpu().numpy()
            else:
                rows = torch.cat((lf0, vuv), dim=1).cpu().numpy()
            off = batch.frame_off.cpu().numpy()
            for u, (name, _) in enumerate(items):
                results[name] = rows[off[u]:off[u + 1]]
        for name in id_list:
            labels = results[name]
            if return_dict:
                label_dict[name] = labels
            if add_deltas:
                if dir_out is not None:
                    labels.tofile(os.path.join(dir_out, LF0LabelGen.dir_deltas, name + LF0LabelGen.ext_deltas))
                ext_deltas.add_sample(labels)
            else:
                if dir_out is not None:
                    labels[:, 0].tofile(os.path.join(dir_out, LF0LabelGen.dir_lf0, name + LF0LabelGen.ext_lf0))
                    labels[:, 1].astype(np.float32).tofile(os.path.join(dir_out, LF0LabelGen.dir_vuv, name + LF0LabelGen.ext_vuv))
                ext_lf0.add_sample(labels[:, :1])
        if not add_deltas:
            if dir_out is not None:
                ext_lf0.save(os.path.join(dir_out, LF0LabelGen.dir_lf0, file_id_list_name))
            norm_lf0 = ext_lf0.get_params()
            norm_first = np.concatenate((norm_lf0[0], (0.0,)), axis=0)
            norm_second = np.concatenate((norm_lf0[1], (1.0,)), axis=0)
        else:
            ext_deltas.sum_frames[-1] = 0.0                          # vuv: mean 0
            ext_deltas.sum_squared_frames[-1] = ext_deltas.sum_length  # variance 1
            if dir_out is not None:
                ext_deltas.save(os.path.join(dir_out, LF0LabelGen.dir_deltas, file_id_list_name))
            norm_first, norm_second = ext_deltas.get_params()
        if return_dict:
            return label_dict, norm_first, norm_second
        return norm_first, norm_second
