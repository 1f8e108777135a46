"""WorldFeatLabelGen with the reference's entry points (idiaptts/src/data_preparation/world/WorldFeatLabelGen.py), computed
by libb200world.so on the GPU:

    convert_to_world_features  :735-762      convert_from_world_features :765-776
    world_extract_features     :779-807      extract_features            :810-889      trim_to_shortest :892-907
    world_features_to_raw      :910-945      gen_data                    :947-1071     save_output      :1121-1172

F0: every extraction entry point takes an optional cached F0 track (`f0=` / `f0_cache=`, north_star); without one the F0
half of pyworld.wav2world (DIO + StoneMask, :792) runs on the device too (ops.estimate_f0), as in the
reference.

gen_data processes the whole id list as ONE ragged GPU batch (the reference loops over utterances, :996) and, when
torch.distributed is initialised, takes this rank's shard of the list and all-reduces the normalisation statistics."""
import glob
import logging
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import corpus_io, distributed, ops, pipeline
from .AudioProcessing import AudioProcessing
from .MeanCovarianceExtractor import MeanCovarianceExtractor
from .MeanStdDevExtractor import MeanStdDevExtractor


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("idiaptts_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class WorldFeatLabelGen(object):
    """Create WORLD feature labels for .wav files."""

    f0_silence_threshold = 30
    lf0_zero = 0
    preemphasis = 0.0
    n_fft = None
    win_length_ms = None

    dir_lf0 = "lf0"
    dir_vuv = "vuv"
    dir_bap = "bap"
    dir_deltas = "cmp"

    ext_lf0 = "lf0"
    ext_vuv = "vuv"
    ext_bap = "bap"
    ext_deltas = "cmp"

    logger = logging.getLogger(__name__)

    class Config(object):
        """WorldFeatLabelGen.Config (world/WorldFeatLabelGen.py:62-138): the reader configuration trainers hand to the data
        pipeline.  Same fields and defaults; the NpzDataReader machinery behind it in the reference is out of scope, so this
        object just carries the fields and builds a reader whose normalisation parameters are loaded (create_reader, :133-138)."""

        class NormType(object):
            NONE = "NONE"
            MEAN_VARIANCE = "MEAN_VARIANCE"
            MEAN_STDDEV = "MEAN_STDDEV"
            MIN_MAX = "MIN_MAX"

        def __init__(self, name, directory=None, indices=None, norm_params_path=None, norm_params=None,
                     norm_type=NormType.MEAN_VARIANCE, output_names=None, preprocessing_fn=None, preprocess_before_norm=False,
                     postprocessing_fn=None, postprocess_before_norm=False, add_deltas=False, preemphasis=0.0, n_fft=None,
                     win_length_ms=None, num_coded_sps=60, num_bap=1, sp_type="mcep", load_sp=True, load_lf0=True, load_vuv=True,
                     load_bap=True, apply_mlpg=True, **kwargs):
            if norm_type is None:
                norm_type = self.NormType.MEAN_VARIANCE if add_deltas else self.NormType.MEAN_STDDEV
            self.name = name
            self.directory = list(directory) if isinstance(directory, (tuple, list)) else [directory]
            self.dir_labels = self.directory[0]
            self.indices = indices
            self.norm_params_path = norm_params_path
            self.norm_params = norm_params
            self.norm_type = norm_type
            self.output_names = output_names if output_names is not None else [name]
            self.preprocessing_fn = preprocessing_fn
            self.preprocess_before_norm = preprocess_before_norm
            self.postprocessing_fn = postprocessing_fn
            self.postprocess_before_norm = postprocess_before_norm
            self.add_deltas = add_deltas
            self.apply_mlpg = apply_mlpg
            self.preemphasis = preemphasis
            self.n_fft = n_fft
            self.win_length_ms = win_length_ms
            self.num_coded_sps = num_coded_sps
            self.num_bap = num_bap
            self.sp_type = sp_type
            self.load_sp, self.load_lf0, self.load_vuv, self.load_bap = load_sp, load_lf0, load_vuv, load_bap
            self.extra = kwargs

        def create_reader(self):
            reader = WorldFeatLabelGen(self)
            if self.norm_params is not None:
                reader.norm_params = self.norm_params
            else:
                reader.get_normalisation_params(self.norm_params_path)
            return reader

    _LEGACY_DEFAULTS = {"add_deltas": False, "preprocessing_fn": None, "preemphasis": 0.0, "n_fft": None, "win_length_ms": None,
                        "num_coded_sps": 60, "num_bap": 1, "sp_type": "mcep", "hop_size_ms": 5, "load_sp": True, "load_lf0": True,
                        "load_vuv": True, "load_bap": True}

    def __init__(self, *args, **kwargs):
        """Two forms, as in the reference (:140-228): WorldFeatLabelGen(config) with a WorldFeatLabelGen.Config, or the legacy
        WorldFeatLabelGen(dir_labels, add_deltas=..., num_coded_sps=..., ...) keyword form.  Extensions of this package (keyword
        only): f0_cache (cached F0 tracks: dict / directory, north_star), mgc_alpha (override of fs_to_mgc_alpha), io_threads
        (host threads of the shard-at-a-time wav / npz file IO, corpus_io.py; 0 = one per hardware thread), io_chunk_seconds (audio
        per piece of gen_data's read / extract / write pipeline), use_distributed (False: gen_data treats the process as the only rank
        even when torch.distributed is initialised -- no sharding, no collective)."""
        self.f0_cache = kwargs.pop("f0_cache", None)
        self.mgc_alpha = kwargs.pop("mgc_alpha", None)
        self.io_threads = int(kwargs.pop("io_threads", 0))
        self.io_chunk_seconds = float(kwargs.pop("io_chunk_seconds", 4096.0))
        self.use_distributed = bool(kwargs.pop("use_distributed", True))
        self._io_slots = None
        if len(args) == 1 and isinstance(args[0], WorldFeatLabelGen.Config):
            config = args[0]
            fields = {k: getattr(config, k) for k in ("add_deltas", "preprocessing_fn", "preemphasis", "n_fft", "win_length_ms",
                                                      "num_coded_sps", "num_bap", "sp_type", "load_sp", "load_lf0", "load_vuv", "load_bap")}
            fields["hop_size_ms"] = 5  # the reference's Config form never sets it; 5 ms is what every recipe uses
            self.dir_labels = config.dir_labels
            self.output_names = config.output_names
            self.apply_mlpg = config.apply_mlpg
            self.legacy_getitem = False
        else:
            if "dir_labels" in kwargs:
                self.dir_labels = kwargs.pop("dir_labels")
            else:
                self.dir_labels = args[0] if args else None
            unknown = set(kwargs) - set(self._LEGACY_DEFAULTS)
            if unknown:
                raise TypeError("unexpected keyword arguments: {}".format(sorted(unknown)))
            fields = dict(self._LEGACY_DEFAULTS)
            fields.update(kwargs)
            self.output_names = ["acoustic_features"]
            self.apply_mlpg = False
            self.legacy_getitem = True
        self.add_deltas = fields["add_deltas"]
        self.preprocessing_fn = fields["preprocessing_fn"]
        self.preemphasis = fields["preemphasis"]
        self.n_fft = fields["n_fft"]
        self.win_length_ms = fields["win_length_ms"]
        self.num_coded_sps = fields["num_coded_sps"]
        self.num_bap = fields["num_bap"]
        self.sp_type = fields["sp_type"]
        self.hop_size_ms = fields["hop_size_ms"]
        self.load_sp, self.load_lf0, self.load_vuv, self.load_bap = (fields["load_sp"], fields["load_lf0"], fields["load_vuv"],
                                                                     fields["load_bap"])
        self.load_flags = (self.load_sp, self.load_lf0, self.load_vuv, self.load_bap)
        self.norm_params = None
        self.covs = [None] * 4  # coded_sp, lf0, (vuv: never used), bap
        self.dir_coded_sps = self.sp_type
        if self.num_coded_sps != -1:
            self.dir_coded_sps += str(self.num_coded_sps)
        self.dir_deltas = WorldFeatLabelGen.dir_deltas + "_" + self.dir_coded_sps
        if self.sp_type not in ("mcep", "mgc"):
            raise NotImplementedError("sp_type '{}': only 'mcep' and 'mgc' are on the accelerated path".format(self.sp_type))

    # ---- normalisation parameters -----------------------------------------------------------------------------------------------
    def get_normalisation_params(self, dir_out=None, file_name=None):
        """WorldFeatLabelGen.get_normalisation_params (:575-732): read the per-feature normalisation files gen_data wrote and set
        self.norm_params = (mean [1, W], std_dev [1, W]) over the loaded features in the order coded_sp, lf0, vuv, bap (vuv: mean 0,
        std_dev 1); with add_deltas also self.covs[idx] (what MLPG needs).  Files per feature sub-directory:
            <file_name->mean-std_dev.npz                     (add_deltas False)
            <file_name->deltas-mean-covariance.npz           (add_deltas True)
        (.bin variants are read when the .npz is missing); then the legacy single-directory layout
        cmp_<sp><D>/<file_name-><feature>-mean-covariance.bin (:661-732).
        Reference quirk not mirrored: with add_deltas AND a file_name the reference raises UnboundLocalError (`full_file_name +=`
        before assignment, :609); here the documented name "<file_name>-deltas" is used."""
        if dir_out is None:
            dir_out = self.dir_labels
        sub_dirs = (self.dir_coded_sps, self.dir_lf0, self.dir_vuv, self.dir_bap)
        has_name = file_name is not None and os.path.basename(file_name) != ""

        def load_any(ext_cls, base):
            try:
                return ext_cls.load(base + ".npz")
            except FileNotFoundError:
                return ext_cls.load(base + ".bin")

        try:
            means, std_devs = [], []
            for idx, (load, subdir) in enumerate(zip(self.load_flags, sub_dirs)):
                if not load:
                    continue
                if subdir == self.dir_vuv:
                    means.append(np.atleast_2d(0.0))
                    std_devs.append(np.atleast_2d(1.0))
                    continue
                prefix = os.path.join(dir_out, subdir, (file_name + "-") if has_name else "")
                if self.add_deltas:
                    mean, cov, std_dev = load_any(MeanCovarianceExtractor, prefix + "deltas-" + MeanCovarianceExtractor.file_name_appendix)
                    self.covs[idx] = cov
                else:
                    mean, std_dev = load_any(MeanStdDevExtractor, prefix + MeanStdDevExtractor.file_name_appendix)
                means.append(np.atleast_2d(mean))
                std_devs.append(np.atleast_2d(std_dev))
            self.norm_params = (np.concatenate(means, axis=1), np.concatenate(std_devs, axis=1))
            return self.norm_params
        except FileNotFoundError as e0:
            # LEGACY layout: one directory cmp_<sp><D> with a mean-covariance .bin per feature (always with deltas)
            prefix = (file_name + "-") if has_name else ""
            means, std_devs = [], []
            for idx, (load, subdir) in enumerate(zip(self.load_flags, sub_dirs)):
                if not load:
                    continue
                if subdir == self.dir_vuv:
                    means.append(np.atleast_2d(0.0))
                    std_devs.append(np.atleast_2d(1.0))
                    continue
                new_style = os.path.join(dir_out, self.dir_deltas, "{}{}-{}.bin".format(prefix, subdir, MeanCovarianceExtractor.file_name_appendix))
                old_style = os.path.join(dir_out, self.dir_deltas, "{}{}_{}.bin".format(prefix, MeanCovarianceExtractor.file_name_appendix, subdir))
                try:
                    mean, cov, std_dev = MeanCovarianceExtractor.load(new_style)
                except FileNotFoundError as e1:
                    try:
                        mean, cov, std_dev = MeanCovarianceExtractor.load(old_style)
                        self.logger.warning("Found legacy style normalisation parameters at %s. Consider recreating features or "
                                            "renaming to %s", old_style, new_style)
                    except FileNotFoundError as e2:
                        raise FileNotFoundError([e0, e1, e2])
                if not self.add_deltas:
                    d = len(cov) // 3
                    assert len(cov) == 3 * d, "Feature size {} is not dividable by 3. Are deltas features contained?".format(len(cov))
                    cov, mean, std_dev = cov[:d, :d], mean[:d], std_dev[:d]
                self.covs[idx] = cov
                means.append(np.atleast_2d(mean))
                std_devs.append(np.atleast_2d(std_dev))
            self.norm_params = (np.concatenate(means, axis=1), np.concatenate(std_devs, axis=1))
            if self.add_deltas:
                self.norm_params = (self.norm_params[0][0], self.norm_params[1][0])
            return self.norm_params

    def _get_norm_params_subset(self, norm_params):
        """:296-330 is a no-op slicing in the reference whenever all four features are loaded; kept for callers that pass
        parameters of ALL features to a reader that loads a subset: returns the columns of the loaded features."""
        mean, std_dev = norm_params
        f3 = 3 if self.add_deltas else 1
        widths = (self.num_coded_sps * f3, f3, 1, self.num_bap * f3)
        keep, pos = [], 0
        for load, w in zip(self.load_flags, widths):
            if load:
                keep.extend(range(pos, pos + w))
            pos += w
        mean, std_dev = np.atleast_2d(mean), np.atleast_2d(std_dev)
        if mean.shape[1] != pos:
            return norm_params
        return mean[:, keep], std_dev[:, keep]

    # ---- layout conversion -------------------------------------------------------------------------------------------
    @staticmethod
    def convert_to_world_features(sample, contains_deltas=False, num_coded_sps=60, num_bap=1):
        deltas_factor = 3 if contains_deltas else 1
        num_expected_feats = (num_coded_sps + 1 + num_bap) * deltas_factor + 1
        if sample.shape[1] != num_expected_feats:
            num_expected_feats = (num_coded_sps + 1 + num_bap) * 3 + 1
            if sample.shape[1] == num_expected_feats:  # deltas detected automatically
                deltas_factor = 3
            else:
                raise ValueError("WORLD requires all features to be present.")
        coded_sp = sample[:, :num_coded_sps]
        lf0 = sample[:, num_coded_sps * deltas_factor]
        vuv = np.copy(sample[:, num_coded_sps * deltas_factor + deltas_factor])
        vuv[vuv < 0.5] = 0.0
        vuv[vuv >= 0.5] = 1.0
        if contains_deltas:
            bap = sample[:, -num_bap * 3:-num_bap * 2]
        else:
            bap = sample[:, -num_bap:]
        return coded_sp, lf0, vuv, bap

    @staticmethod
    def convert_from_world_features(coded_sp, lf0, vuv, bap):
        if lf0.ndim < 2:
            lf0 = lf0[:, None]
        if vuv.ndim < 2:
            vuv = vuv[:, None]
        if bap.ndim < 2:
            bap = bap[:, None]
        return np.concatenate((coded_sp, lf0, vuv, bap), axis=1)

    # ---- F0 cache ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _lookup_f0(f0_cache, file_name):
        """f0_cache: dict {id: f0 array} | directory holding <id>.npy (float64 Hz per 5 ms frame, 0 = unvoiced) | None."""
        key = os.path.basename(file_name)
        if f0_cache is None:
            return None  # no cache: DIO + StoneMask on the device, as pyworld.wav2world does (:792)
        if isinstance(f0_cache, dict):
            if key in f0_cache:
                return np.asarray(f0_cache[key], np.float64)
            if file_name in f0_cache:
                return np.asarray(f0_cache[file_name], np.float64)
            raise KeyError("no cached F0 for %s" % file_name)
        path = os.path.join(f0_cache, key + ".npy")
        return np.load(path).astype(np.float64)

    # ---- analysis -------------------------------------------------------------------------------------------------------
    @staticmethod
    def world_extract_features(raw, fs, hop_size_ms, f0_silence_threshold=None, lf0_zero=None, n_fft=None, f0=None):
        """Returns (amp_sp [T, K] float64, lf0 [T, 1] float32, vuv [T, 1] float32, bap [T, nap] float32)."""
        if f0_silence_threshold is None:
            f0_silence_threshold = WorldFeatLabelGen.f0_silence_threshold
        if lf0_zero is None:
            lf0_zero = WorldFeatLabelGen.lf0_zero
        raw = np.ascontiguousarray(raw)
        if raw.dtype not in (np.float64, np.float32, np.int16):
            raw = raw.astype(np.float64)
        T = ops.num_frames(len(raw), fs, hop_size_ms)
        estimate = f0 is None
        f0 = np.zeros(T) if estimate else np.ascontiguousarray(f0, np.float64)
        if len(f0) != T:
            raise ValueError("cached F0 has {} frames, the waveform gives {} at {} ms hop".format(len(f0), T, hop_size_ms))
        dev = _device()
        fft_size = n_fft if n_fft is not None else ops.get_cheaptrick_fft_size(fs)
        batch = ops.RaggedBatch.from_host([raw], [f0], fs, frame_period=hop_size_ms, device=dev)
        if estimate:
            ops.estimate_f0(batch, frame_period=hop_size_ms)
        status = ops.new_status(dev)
        sp, _ = ops.cheaptrick(batch, fft_size=fft_size, out_dtype=torch.float64, status=status)
        amp_sp = torch.sqrt(sp)
        coarse, voiced, _ = ops.d4c_coarse(batch, status=status)
        bap = ops.bap_from_coarse(coarse, voiced, fs, fft_size)
        lf0, vuv = ops.lf0_vuv(batch.f0, batch.frame_off, f0_silence_threshold, lf0_zero)
        out = (amp_sp.cpu().numpy(), lf0.cpu().numpy(), vuv.cpu().numpy(), bap.cpu().numpy())
        ops.raise_for_status(status, "world_extract_features")
        return out

    @staticmethod
    def extract_features(dir_in, file_name, file_ext="wav", preemphasis=0.0, n_fft=None, win_length_ms=None, hop_size_ms=5,
                         sp_type="mcep", num_coded_sps=40, load_sp=True, load_lf0=True, load_vuv=True, load_bap=True,
                         f0_silence_threshold=None, lf0_zero=None, f0=None, f0_cache=None, mgc_alpha=None):
        """Acoustic features of one audio file: (coded_sp, lf0, vuv, bap)."""
        if sp_type not in ("mcep", "mgc"):
            raise NotImplementedError("sp_type '{}': only 'mcep' and 'mgc' are on the accelerated path".format(sp_type))
        audio_name = os.path.join(dir_in, file_name + "." + file_ext)
        raw, fs = AudioProcessing.get_raw(audio_name, preemphasis)
        if f0 is None:
            f0 = WorldFeatLabelGen._lookup_f0(f0_cache, file_name)
        amp_sp, lf0, vuv, bap = WorldFeatLabelGen.world_extract_features(raw, fs, hop_size_ms, f0_silence_threshold, lf0_zero,
                                                                         n_fft, f0=f0)
        if load_vuv:
            voiced_pct = vuv.sum() / len(vuv) * 100.0
            if voiced_pct < 5.0:
                logging.warning("Detected only {:.0f}% [{}/{}] unvoiced frames in {}.".format(voiced_pct, int(vuv.sum()),
                                                                                              len(vuv), file_name))
        coded_sp = None
        if load_sp:
            a = AudioProcessing.fs_to_mgc_alpha(fs) if mgc_alpha is None else mgc_alpha
            if sp_type == "mcep":
                coded_sp = AudioProcessing.extract_mcep(amp_sp, num_coded_sps=num_coded_sps, mgc_alpha=a)
            else:
                coded_sp = AudioProcessing.extract_mgc(amp_sp, fs=fs, num_coded_sps=num_coded_sps, mgc_alpha=a)
            assert len(coded_sp) == len(lf0), "Requires testing. Possibly trimming is a solution."
        logging.info("Extracted features from {} at {} Hz with {} ms frame hop.".format(os.path.basename(file_name), fs, hop_size_ms))
        coded_sp, lf0, vuv, bap = WorldFeatLabelGen.trim_to_shortest([coded_sp, lf0, vuv, bap])
        return coded_sp, lf0, vuv, bap

    @staticmethod
    def trim_to_shortest(features):
        len_shortest = min(map(len, [f for f in features if f is not None]))
        for idx, feature in enumerate(features):
            if feature is None:
                continue
            len_diff = len(feature) - len_shortest
            if len_diff > 0:
                trim_front = len_diff // 2
                trim_end = len_diff - trim_front
                features[idx] = feature[trim_front:len(feature) - trim_end]
        return features

    # ---- synthesis ------------------------------------------------------------------------------------------------------
    @staticmethod
    def world_features_to_raw(amp_sp, lf0, vuv, bap, fs, n_fft=None, f0_silence_threshold=None, lf0_zero=None, preemphasis=0.0):
        """WORLD synthesis of one utterance -> waveform (float32 values; float64 array as the reference returns after
        scipy's lfilter)."""
        if f0_silence_threshold is None:
            f0_silence_threshold = WorldFeatLabelGen.f0_silence_threshold
        if lf0_zero is None:
            lf0_zero = WorldFeatLabelGen.lf0_zero
        if n_fft is None:
            n_fft = AudioProcessing.fs_to_frame_length(fs)
        dev = _device()
        pow_sp = np.square(amp_sp, dtype=np.float64)
        f0 = np.exp(lf0, dtype=np.float64)
        vuv = np.array(vuv, copy=True)
        vuv[f0 < f0_silence_threshold] = 0
        f0[vuv == 0] = lf0_zero
        if f0.ndim > 1:
            assert f0.shape[1:] == (1,) * (f0.ndim - 1), "F0 should have only one dimension at this stage."
            f0 = f0.squeeze()
        if bap.ndim < 2:
            bap = bap.reshape(-1, 1)
        ap = ops.decode_aperiodicity(torch.from_numpy(np.ascontiguousarray(bap, np.float64)).to(dev), fs, n_fft)
        frame_off = torch.tensor([0, len(f0)], dtype=torch.int64, device=dev)
        y, _, status = ops.synthesize(torch.from_numpy(np.ascontiguousarray(f0)).to(dev),
                                      torch.from_numpy(np.ascontiguousarray(pow_sp)).to(dev), ap, frame_off, fs,
                                      deemphasis=0.0, out_dtype=torch.float32)
        raw = y.cpu().numpy()
        ops.raise_for_status(status, "world_features_to_raw")
        return AudioProcessing.depreemphasis(raw, preemphasis)

    # ---- corpus extraction -------------------------------------------------------------------------------------------------
    def gen_data(self, dir_in, dir_out=None, file_id_list="", file_ext="wav", id_list=None, return_dict=False, f0_cache=None):
        """Prepare acoustic features of all utterances in id_list in ragged GPU batches (pieces of io_chunk_seconds of audio, read,
        extracted and written by a three-stage pipeline: _extract_shard); save them per feature / utterance as .npz; return
        ([label_dict,] means, std_devs).  With torch.distributed initialised each rank extracts its shard and the statistics are
        all-reduced (files are written by the owning rank, statistics by rank 0)."""
        id_list, file_id_list_name = self._get_id_list(dir_in, file_id_list, id_list, file_ext)
        f0_cache = f0_cache if f0_cache is not None else self.f0_cache
        rank, world = distributed.world_info() if self.use_distributed else (0, 1)
        if dir_out is not None:
            self._create_directories(dir_out)
        dev = _device()

        # read this rank's shard
        wav_paths = [os.path.join(dir_in, name + "." + file_ext) for name in id_list]
        info = corpus_io.probe_wavs(wav_paths, threads=self.io_threads)  # all RIFF headers in one native call
        headers = list(zip(info["num_samples"].tolist(), info["fs"].tolist()))
        if world > 1:
            mine = distributed.shard_utterances([h[0] for h in headers], world)[rank]
        else:
            mine = np.arange(len(id_list))
        my_ids = [id_list[i] for i in mine]
        # every rank sizes the statistics buffer from the wav HEADERS (all ranks read all headers), so a rank whose shard fails
        # still takes part in the all-reduce: the failure travels in the last element and every rank raises after the collective
        fs = headers[0][1] if headers else 16000
        nap = ops.get_num_aperiodicities(fs)
        D = self.num_coded_sps
        dim = D + 2 + nap
        alpha = self.mgc_alpha if self.mgc_alpha is not None else AudioProcessing.fs_to_mgc_alpha(fs)
        stat = torch.zeros(1 + 2 * dim * (3 if self.add_deltas else 1) + (3 * dim) ** 2 * (1 if self.add_deltas else 0) + 1,
                           dtype=torch.float64, device=dev)
        # per-feature column groups of the static block
        groups = (("sp", self.load_sp, self.dir_coded_sps, self.sp_type, slice(0, D)),
                  ("lf0", self.load_lf0, self.dir_lf0, self.ext_lf0, slice(D, D + 1)),
                  ("vuv", self.load_vuv, self.dir_vuv, self.ext_vuv, slice(D + 1, D + 2)),
                  ("bap", self.load_bap, self.dir_bap, self.ext_bap, slice(D + 2, D + 2 + nap)))

        def cols(sl, block):  # columns of a feature in block 0 (static), 1 (delta), 2 (delta-delta)
            return np.arange(sl.start, sl.stop) + block * dim

        label_dict = OrderedDict()
        failure = None
        try:
            if np.any(info["fs"][mine] != fs):
                raise ValueError("mixed sampling rates in one gen_data call ({} vs {})".format(
                    fs, int(info["fs"][mine][info["fs"][mine] != fs][0])))
            if len(my_ids):
                self._extract_shard(dev, wav_paths, info, mine, my_ids, f0_cache, fs, alpha, dim, stat, groups, cols, dir_out,
                                    label_dict if return_dict else None)
        except Exception as e:  # noqa: BLE001 -- re-raised below, after the collective
            failure = e
            stat.zero_()
            stat[-1] = 1.0
        if world > 1:
            distributed.allreduce_stats(stat)
        stat_np = stat.cpu().numpy()
        if failure is not None:
            raise failure
        if stat_np[-1] > 0:
            raise RuntimeError("gen_data failed on {} other rank(s); see their error".format(int(round(stat_np[-1]))))
        stat_np = stat_np[:-1]
        n_total = int(round(stat_np[0]))

        # normalisation parameters from the (all-reduced) sums
        output_means, output_std_dev = [], []
        for key, load, fdir, fext, sl in groups:
            if not load:
                continue
            if key == "vuv":
                mean, std = np.atleast_1d(0.0), np.atleast_1d(1.0)
                output_means.append(mean)
                output_std_dev.append(std)
                continue
            if self.add_deltas:
                c = np.concatenate([cols(sl, b) for b in range(3)])
                ext = MeanCovarianceExtractor()
                gram = stat_np[1 + 6 * dim:].reshape(3 * dim, 3 * dim)
                ext.add_sums(n_total, stat_np[1:1 + 3 * dim][c][None, :], gram[np.ix_(c, c)])
                mean, cov = ext.get_params()
                output_means.append(mean)
                output_std_dev.append(cov)
            else:
                ext = MeanStdDevExtractor()
                ext.add_sums(n_total, stat_np[1:1 + dim][sl], stat_np[1 + dim:1 + 2 * dim][sl])
                mean, std = ext.get_params()
                output_means.append(mean)
                output_std_dev.append(std)
            if dir_out and rank == 0:
                norm_file_path = os.path.join(dir_out, fdir, file_id_list_name)
                if self.add_deltas:
                    if file_id_list_name is not None and os.path.basename(file_id_list_name) != "":
                        norm_file_path += "-"
                    norm_file_path += "deltas"
                self.logger.info("Write norm_prams to {}".format(norm_file_path))
                ext.save(norm_file_path)
        if not self.add_deltas:
            if len(output_means) > 0:
                output_means = np.concatenate(output_means, axis=0)
                output_std_dev = np.concatenate(output_std_dev, axis=0)
            else:
                output_means, output_std_dev = None, None
        self.norm_params = (output_means, output_std_dev)
        if self.add_deltas:  # covariance matrices per LOADED feature -> the reference's indexing (coded_sp, lf0, vuv, bap)
            loaded = [i for i, l in enumerate(self.load_flags) if l]
            self.covs = [None] * 4
            for i, c in zip(loaded, output_std_dev):
                self.covs[i] = None if i == 2 else np.asarray(c, np.float32)
        if return_dict:
            return label_dict, output_means, output_std_dev
        return output_means, output_std_dev

    def _extract_shard(self, dev, wav_paths, info, mine, my_ids, f0_cache, fs, alpha, dim, stat, groups, cols, dir_out, label_dict):
        """The body of gen_data for this rank's shard, as a three-stage pipeline over pieces of about io_chunk_seconds of audio:
            reader thread   wav files -> one pinned int16 buffer (native thread pool, corpus_io.read_wavs_i16)
            this thread     pinned buffer -> device, the analysis kernels, statistics, feature rows -> pinned buffer (copy stream)
            writer thread   pinned feature rows -> per-feature .npz archives (corpus_io.write_npz) [and the label dictionary]
        with two buffers per hand-over, so the file reads of piece k + 1 and the file writes of piece k - 1 run while the GPU works
        on piece k, and host memory is bounded by the piece size instead of the corpus.  The statistics accumulate on the device
        across pieces (`stat` layout: [N | sums | (gram)] + failure flag)."""
        import queue
        import threading
        D = self.num_coded_sps
        width = dim * (3 if self.add_deltas else 1)
        packed_ok = bool(np.all(info["bits"][mine] == 16) and np.all(info["channels"][mine] == 1))
        # pieces of whole utterances, about io_chunk_seconds of audio each
        budget = int(self.io_chunk_seconds * fs)
        pieces, cur, acc = [], [], 0
        for j, i in enumerate(mine):
            n = int(info["num_samples"][i])
            if cur and acc + n > budget:
                pieces.append(cur)
                cur, acc = [], 0
            cur.append(j)
            acc += n
        if cur:
            pieces.append(cur)
        frames_of = [ops.num_frames(int(info["num_samples"][i]), fs, self.hop_size_ms) for i in mine]
        max_samples = max(sum(int(info["num_samples"][mine[j]]) for j in pc) for pc in pieces)
        max_frames = max(sum(frames_of[j] for j in pc) for pc in pieces)
        want_rows = dir_out is not None or label_dict is not None
        nslots = min(2, len(pieces))
        slots = self._io_slots
        if (slots is None or slots["in"][0].numel() < max_samples or slots["out"][0].shape[0] < max_frames
                or slots["out"][0].shape[1] != width or len(slots["in"]) < nslots):
            slots = {"in": [torch.empty(max_samples if packed_ok else 0, dtype=torch.int16, pin_memory=packed_ok and max_samples > 0)
                            for _ in range(nslots)],
                     "out": [torch.empty((max_frames if want_rows else 0, width), dtype=torch.float32,
                                         pin_memory=want_rows and max_frames > 0) for _ in range(nslots)]}
            self._io_slots = slots  # pinned allocations are expensive: kept for the next call
        an = pipeline.WorldAnalyzer(fs, D, alpha, self.hop_size_ms, self.n_fft, WorldFeatLabelGen.f0_silence_threshold,
                                    WorldFeatLabelGen.lf0_zero, device=dev, sp_type=self.sp_type, mgc_gamma=AudioProcessing.mgc_gamma)
        status = ops.new_status(dev)
        sums = torch.zeros(2 * dim, dtype=torch.float64, device=dev)
        sums3 = torch.zeros(2 * 3 * dim, dtype=torch.float64, device=dev) if self.add_deltas else None
        gram = torch.zeros((3 * dim) ** 2, dtype=torch.float64, device=dev) if self.add_deltas else None
        copy_stream = torch.cuda.Stream(dev)
        cur_stream = torch.cuda.current_stream(dev)
        free_in, ready_in, free_out, ready_out = queue.Queue(), queue.Queue(), queue.Queue(), queue.Queue()
        for k in range(nslots):
            free_in.put((k, None))
            free_out.put(k)
        errors = []
        stop = threading.Event()

        def reader():
            try:
                for pc in pieces:
                    if stop.is_set():
                        break
                    f0s = []
                    for j in pc:
                        name = my_ids[j]
                        f0 = self._lookup_f0(f0_cache, name)
                        if f0 is None:
                            f0 = np.zeros(frames_of[j])
                        if len(f0) != frames_of[j]:
                            raise ValueError("{}: cached F0 has {} frames, the waveform gives {}".format(name, len(f0), frames_of[j]))
                        f0s.append(f0)
                    paths = [wav_paths[mine[j]] for j in pc]
                    if packed_ok:
                        slot, ev = free_in.get()
                        if ev is not None:
                            ev.synchronize()  # the previous piece in this buffer has reached the device
                        idx = np.asarray([mine[j] for j in pc])
                        sample_off = np.concatenate(([0], np.cumsum(info["num_samples"][idx]))).astype(np.int64)
                        corpus_io._read_into(paths, np.ascontiguousarray(info["data_offset"][idx]), sample_off, slots["in"][slot],
                                             self.io_threads)
                        ready_in.put((pc, f0s, slot, sample_off, None))
                    else:  # other sample widths: per file, as float64
                        waves = [AudioProcessing.read_wav(p)[0] for p in paths]
                        waves = [w.astype(np.float64) / 32768.0 if w.dtype == np.int16 else w.astype(np.float64) for w in waves]
                        ready_in.put((pc, f0s, None, None, waves))
            except Exception as e:  # noqa: BLE001
                errors.append(e)
            finally:
                ready_in.put(None)

        sel = None
        if label_dict is not None:
            sel = np.concatenate([np.concatenate([cols(sl, blk) for blk in range(3 if self.add_deltas and key != "vuv" else 1)])
                                  for key, load, fdir, fext, sl in groups if load] or [np.zeros(0, np.int64)]).astype(np.int64)

        def writer():
            while True:
                item = ready_out.get()
                if item is None:
                    break
                pc, slot, ev, offs = item
                try:
                    if errors:
                        continue
                    ev.synchronize()
                    rows = slots["out"][slot]
                    if dir_out is not None:
                        # one native call per feature directory writes the archives of the piece (LabelGen.save_output's layout:
                        # <dir>/<feat>/<id>.npz with key <ext>[, <ext>_deltas, <ext>_double_deltas])
                        bases = [os.path.basename(my_ids[j]) for j in pc]
                        for key, load, fdir, fext, sl in groups:
                            if not load:
                                continue
                            with_deltas = self.add_deltas and key != "vuv"
                            keys = [fext, fext + "_deltas", fext + "_double_deltas"] if with_deltas else [fext]
                            corpus_io.write_npz([os.path.join(dir_out, fdir, b + ".npz") for b in bases], keys,
                                                [sl.start + blk * dim for blk in range(len(keys))], [sl.stop - sl.start] * len(keys),
                                                offs, rows, threads=self.io_threads)
                    if label_dict is not None:
                        rows_np = rows.numpy()
                        for u, j in enumerate(pc):
                            label_dict[my_ids[j]] = rows_np[offs[u]:offs[u + 1]][:, sel] if len(sel) else None  # fancy index: a copy
                except Exception as e:  # noqa: BLE001
                    errors.append(e)
                    stop.set()
                finally:
                    free_out.put(slot)  # always handed back: the extraction loop may be waiting for it

        threads = [threading.Thread(target=reader, name="b2w-wav-reader", daemon=True)]
        if want_rows:
            threads.append(threading.Thread(target=writer, name="b2w-npz-writer", daemon=True))
        for t in threads:
            t.start()
        total_frames = 0
        try:
            while True:
                item = ready_in.get()
                if item is None or errors:
                    break
                pc, f0s, slot, sample_off, waves = item
                if slot is not None:
                    batch = ops.RaggedBatch.from_packed(slots["in"][slot][:int(sample_off[-1])], sample_off, f0s, fs,
                                                        frame_period=self.hop_size_ms, preemphasis=self.preemphasis, device=dev)
                    ev = torch.cuda.Event()
                    ev.record(cur_stream)
                    free_in.put((slot, ev))
                else:
                    batch = ops.RaggedBatch.from_host(waves, f0s, fs, frame_period=self.hop_size_ms, preemphasis=self.preemphasis,
                                                      device=dev)
                if f0_cache is None:
                    ops.estimate_f0(batch, frame_period=self.hop_size_ms)
                feats, _, _ = an.extract(batch, sums=sums, status=status)
                F = batch.num_frames
                total_frames += F
                if self.add_deltas:
                    d, dd = ops.deltas(feats, batch.frame_off)
                    feats = torch.cat((feats, d, dd), dim=1).contiguous()  # [F, 3*dim] = [static | delta | delta-delta]
                    ops.stats_accumulate(feats, sums3, gram)
                if want_rows:
                    oslot = free_out.get()  # blocks while the writer still holds both buffers
                    done = torch.cuda.Event()
                    done.record(cur_stream)
                    copy_stream.wait_event(done)
                    with torch.cuda.stream(copy_stream):
                        slots["out"][oslot][:F].copy_(feats, non_blocking=True)
                        copied = torch.cuda.Event()
                        copied.record(copy_stream)
                    feats.record_stream(copy_stream)
                    offs = np.concatenate(([0], np.cumsum([frames_of[j] for j in pc]))).astype(np.int64)
                    ready_out.put((pc, oslot, copied, offs))
        finally:
            stop.set()
            if want_rows:
                ready_out.put(None)
            # unblock a reader that waits for a buffer, then collect the threads
            free_in.put((0, None))
            free_in.put((1, None))
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        if self.add_deltas:
            stat[1:1 + 6 * dim] = sums3
            stat[1 + 6 * dim:-1] = gram
        else:
            stat[1:1 + 2 * dim] = sums
        stat[0] = float(total_frames)
        ops.raise_for_status(status, "gen_data")

    # ---- reading features back (the on-disk formats at the boundary) ----------------------------------------
    @staticmethod
    def load_sample(id_name, dir_out, add_deltas=False, num_coded_sps=60, num_bap=1, sp_type="mcep", load_sp=True, load_lf0=True,
                    load_vuv=True, load_bap=True):
        """WorldFeatLabelGen.load_sample (world/WorldFeatLabelGen.py:418-457): [T, len(coded_sp, lf0, vuv, bap)] float32 from
        the per-feature npz files gen_data wrote (no pre-processing)."""
        assert dir_out is not None, "dir_out cannot be None"
        id_name = os.path.splitext(os.path.basename(id_name))[0]
        reader = WorldFeatLabelGen(dir_labels=dir_out, add_deltas=add_deltas, num_coded_sps=num_coded_sps, num_bap=num_bap,
                                   sp_type=sp_type, load_sp=load_sp, load_lf0=load_lf0, load_vuv=load_vuv, load_bap=load_bap)
        return reader.load(id_name)

    def load(self, id_name):
        """:459-567: `<dir>/<feat>/<id>.npz` with keys <ext>[, <ext>_deltas, <ext>_double_deltas]; falls back to the legacy
        raw-float32 files `<id>.<ext>` (no deltas) and `cmp_<sp><D>/<id>.cmp` (deltas)."""
        f3 = 3 if self.add_deltas else 1
        feats = (("sp", self.load_sp, self.dir_coded_sps, self.sp_type, self.num_coded_sps * f3),
                 ("lf0", self.load_lf0, self.dir_lf0, self.ext_lf0, f3),
                 ("vuv", self.load_vuv, self.dir_vuv, self.ext_vuv, 1),
                 ("bap", self.load_bap, self.dir_bap, self.ext_bap, self.num_bap * f3))
        out = []
        try:
            for key, load, fdir, fext, fdim in feats:
                if not load:
                    continue
                path = os.path.join(self.dir_labels, fdir, id_name)
                if os.path.exists(path + ".npz"):
                    with np.load(path + ".npz") as arc:
                        lab = arc[fext]
                        if self.add_deltas and key != "vuv":
                            lab = np.concatenate((lab, arc[fext + "_deltas"], arc[fext + "_double_deltas"]), axis=1)
                elif not self.add_deltas:
                    lab = np.fromfile(path + "." + fext, dtype=np.float32).reshape(-1, fdim)  # LEGACY raw float32
                else:
                    raise FileNotFoundError(path + ".npz")
                out.append(lab)
        except FileNotFoundError:
            if not self.add_deltas:
                raise
            # LEGACY cmp file: [sp d dd | lf0 d dd | vuv | bap d dd]
            D, nap = self.num_coded_sps, self.num_bap
            path = os.path.join(self.dir_labels, "{}_{}{}".format(WorldFeatLabelGen.dir_deltas, self.sp_type, D),
                                "{}.{}".format(id_name, WorldFeatLabelGen.ext_deltas))
            lab = np.fromfile(path, dtype=np.float32).reshape(-1, 3 * (D + 1 + nap) + 1)
            out = []
            if self.load_sp:
                out.append(lab[:, :3 * D])
            if self.load_lf0:
                out.append(lab[:, 3 * D:3 * D + 3])
            if self.load_vuv:
                out.append(lab[:, -3 * nap - 1:-3 * nap])
            if self.load_bap:
                out.append(lab[:, -3 * nap:])
        return np.concatenate(out, axis=1)

    def load_batch(self, id_names, pin=True, verify_crc=True):
        """The reader protocol for a whole mini-batch / shard: the rows `load` returns for every id, packed -- (feats, frame_off)
        with feats a float32 CPU torch tensor [F, width] (pinned on request: the next stop is prepare_batch_device) and frame_off
        int64 [U + 1].  One native call per feature directory reads all archives with a pool of threads (corpus_io.read_npz).
        Archives the native reader refuses (compressed, other dtypes) and the legacy raw formats go through `load` per id."""
        ids = [os.path.splitext(os.path.basename(i))[0] for i in id_names]
        f3 = 3 if self.add_deltas else 1
        feats_def = (("sp", self.load_sp, self.dir_coded_sps, self.sp_type, self.num_coded_sps),
                     ("lf0", self.load_lf0, self.dir_lf0, self.ext_lf0, 1),
                     ("vuv", self.load_vuv, self.dir_vuv, self.ext_vuv, 1),
                     ("bap", self.load_bap, self.dir_bap, self.ext_bap, self.num_bap))
        plan, width = [], 0
        for key, load, fdir, fext, d in feats_def:
            if not load:
                continue
            nblk = f3 if key != "vuv" else 1
            keys = [fext, fext + "_deltas", fext + "_double_deltas"][:nblk]
            plan.append((fdir, keys, [width + b * d for b in range(nblk)], [d] * nblk))
            width += nblk * d
        try:
            if not plan or not ids:
                raise ValueError("nothing to read natively")
            paths0 = [os.path.join(self.dir_labels, plan[0][0], i + ".npz") for i in ids]
            rows, _ = corpus_io.probe_npz(paths0, plan[0][1][0], threads=self.io_threads)
            frame_off = np.concatenate(([0], np.cumsum(rows))).astype(np.int64)
            feats = torch.empty((int(frame_off[-1]), width), dtype=torch.float32, pin_memory=bool(pin and frame_off[-1] > 0))
            for fdir, keys, col_off, cols in plan:
                corpus_io.read_npz([os.path.join(self.dir_labels, fdir, i + ".npz") for i in ids], keys, col_off, cols, frame_off,
                                   feats, verify_crc=verify_crc, threads=self.io_threads)
            return feats, frame_off
        except ValueError:
            parts = [self.load(i) for i in ids]
            frame_off = np.concatenate(([0], np.cumsum([len(p) for p in parts]))).astype(np.int64)
            feats = torch.from_numpy(np.ascontiguousarray(np.concatenate(parts, axis=0), np.float32)) if parts else torch.zeros((0, width))
            return (feats.pin_memory() if pin and feats.numel() else feats), frame_off

    def __getitem__(self, id_name):
        """Load and normalise one sample (the reference's reader protocol, :290-300): the legacy constructor form returns the
        array, the Config form a dict {output_name: array}.  Normalisation parameters must have been set
        (get_normalisation_params / gen_data / Config.norm_params): a reader without them raises instead of silently
        returning un-normalised features."""
        if self.norm_params is None:
            raise RuntimeError("normalisation parameters are not loaded: call get_normalisation_params(dir_out, file_name) first")
        sample = self.load(os.path.splitext(os.path.basename(id_name))[0])
        if self.preprocessing_fn is not None:
            sample = self.preprocessing_fn(sample)
        sample = self.preprocess_sample(sample)
        return sample if self.legacy_getitem else {self.output_names[0]: sample}

    def _flat_norm_params(self, norm_params=None):
        mean, std = norm_params if norm_params is not None else (self.norm_params if self.norm_params is not None else (None, None))
        if mean is None:
            return None, None
        if isinstance(mean, (list, tuple)):  # add_deltas: per-feature means / covariance matrices
            mean = np.concatenate([np.atleast_1d(m).reshape(-1) for m in mean])
            std = np.concatenate([np.sqrt(np.diag(c)) if np.ndim(c) == 2 else np.atleast_1d(c).reshape(-1) for c in std])
        return np.asarray(mean, np.float32), np.asarray(std, np.float32)

    def preprocess_sample(self, sample, norm_params=None):
        """(sample - mean) / std_dev with the parameters of gen_data / get_normalisation_params (vuv: mean 0, std 1)."""
        mean, std = self._flat_norm_params(norm_params)
        if mean is None:
            return sample
        return np.float32((sample - mean) / std)

    def prepare_batch_device(self, feats, frame_off, batch_first=False, norm_params=None, min_frames=None, lengths=None):
        """Device-side form of `preprocess_sample` on every utterance + `ModularModelHandlerPyTorch.prepare_batch` (:389-499) for
        feature rows that are already on the GPU (e.g. straight out of pipeline.WorldAnalyzer.extract): ragged [F, W] float32 +
        frame offsets -> (padded [T_max, B, W] | [B, T_max, W], mask, lengths).  One HBM pass (ops.pad_normalise)."""
        mean, std = self._flat_norm_params(norm_params)
        dev = feats.device
        m = None if mean is None else torch.from_numpy(np.ascontiguousarray(mean.reshape(-1))).to(dev)
        sd = None if std is None else torch.from_numpy(np.ascontiguousarray(std.reshape(-1))).to(dev)
        return ops.pad_normalise(feats, frame_off, m, sd, batch_first=batch_first, min_frames=min_frames, lengths=lengths)

    def batches(self, id_list, batch_size, shuffle=False, seed=0, batch_first=False, drop_last=False, prefetch=2, device=None,
                norm_params=None):
        """The trainer-facing loader for this reader (the role of the reference's DataLoader over `__getitem__` +
        `ModularModelHandlerPyTorch.prepare_batch`, :389-499): yields, per mini-batch, a dict with `ids`, `padded`
        ([T_max, B, W] or [B, T_max, W] float32 on the device, normalised), `mask`, `lengths` (host int64 [B]) and `frame_off`
        (device int64 [B + 1]).  A background thread reads the archives of the next `prefetch` mini-batches into pinned host
        memory (load_batch: one native call per feature directory) while the caller trains on the current one; upload and
        pad + normalise (one HBM pass) happen on the caller's stream.  Normalisation parameters must be loaded."""
        import queue
        import threading
        if self.norm_params is None and norm_params is None:
            raise RuntimeError("normalisation parameters are not loaded: call get_normalisation_params(dir_out, file_name) first")
        dev = torch.device(device) if device is not None else _device()
        ids = list(id_list)
        if shuffle:
            order = np.random.default_rng(seed).permutation(len(ids))
            ids = [ids[i] for i in order]
        groups = [ids[i:i + batch_size] for i in range(0, len(ids), batch_size)]
        if drop_last and groups and len(groups[-1]) < batch_size:
            groups.pop()
        q = queue.Queue(maxsize=max(1, int(prefetch)))
        stop = threading.Event()

        def put(item):  # a bounded put that gives up when the consumer has gone away
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def worker():
            try:
                for g in groups:
                    if stop.is_set() or not put((g,) + self.load_batch(g, pin=dev.type == "cuda")):
                        return
                put(None)
            except Exception as e:  # noqa: BLE001 -- re-raised in the consumer
                put(e)

        t = threading.Thread(target=worker, name="b2w-batch-loader", daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, Exception):
                    raise item
                g, feats, frame_off = item
                fo_dev = torch.from_numpy(frame_off).to(dev)
                padded, mask, lengths = self.prepare_batch_device(feats.to(dev, non_blocking=True), fo_dev, batch_first=batch_first,
                                                                  norm_params=norm_params, lengths=np.diff(frame_off))
                yield {"ids": g, "padded": padded, "mask": mask, "lengths": lengths, "frame_off": fo_dev}
        finally:
            stop.set()
            t.join()

    def unprepare_batch_device(self, padded, frame_off, frame_utt, batch_first=False, norm_params=None):
        """Network output [T_max, B, W] | [B, T_max, W] -> de-normalised ragged rows [F, W] on the device (the first half of
        `postprocess_sample`, :338-355; MLPG / synthesis continue from there without leaving the GPU)."""
        mean, std = self._flat_norm_params(norm_params)
        dev = padded.device
        m = None if mean is None else torch.from_numpy(np.ascontiguousarray(mean.reshape(-1))).to(dev)
        sd = None if std is None else torch.from_numpy(np.ascontiguousarray(std.reshape(-1))).to(dev)
        return ops.unpad_denormalise(padded, frame_off, frame_utt, m, sd, batch_first=batch_first)

    def postprocess_sample(self, sample, norm_params=None, apply_mlpg=True):
        """:338-355: de-normalise, then _postprocess_world (MLPG per feature when the sample carries deltas)."""
        mean, std = self._flat_norm_params(norm_params)
        if mean is not None:
            sample = np.copy((sample * std) + mean)
        return self._postprocess_world(sample, apply_mlpg=apply_mlpg)

    # ---- after inference: [static | delta | delta-delta] network output -> WORLD features (MLPG) ------------------------------
    def _postprocess_world(self, sample, norm_params=None, apply_mlpg=True):
        """WorldFeatLabelGen._postprocess_world (world/WorldFeatLabelGen.py:357-415): with add_deltas the (already de-normalised)
        sample [T, 3 D_sp + 3 + 1 + 3 nap] is reduced to [coded_sp | lf0 | vuv | bap] by MLPG per feature (self.covs[0], [1], [3]:
        the covariance matrices gen_data returns with add_deltas) or, with apply_mlpg=False, by keeping the static columns;
        vuv is thresholded at 0.5.  All features of the call go through the CUDA MLPG solver (idiaptts_b200.mlpg)."""
        if not self.add_deltas:
            return sample
        from .mlpg import MLPG
        mlpg = MLPG()
        output_list = list()
        num_processed_feats = 0
        if self.load_sp:
            coded_sp_full = sample[:, :self.num_coded_sps * 3]
            num_processed_feats += self.num_coded_sps * 3
            if apply_mlpg:
                coded_sp = mlpg.generation(coded_sp_full, self.covs[0], self.covs[0].shape[0] // 3)
            else:
                coded_sp = coded_sp_full[:, :self.num_coded_sps]
            output_list.append(coded_sp)
        if self.load_lf0:
            lf0_full = sample[:, num_processed_feats:num_processed_feats + 3]
            num_processed_feats += 3
            if apply_mlpg:
                lf0 = mlpg.generation(lf0_full, self.covs[1], self.covs[1].shape[0] // 3)
            else:
                lf0 = lf0_full[:, 0:1]
            output_list.append(lf0)
        if self.load_vuv:
            vuv = sample[:, num_processed_feats]
            num_processed_feats += 1
            vuv[vuv <= 0.5] = 0.0
            vuv[vuv > 0.5] = 1.0
            output_list.append(vuv[:, None])
        if self.load_bap:
            bap_full = sample[:, -self.num_bap * 3:]
            if apply_mlpg:
                bap = mlpg.generation(bap_full, self.covs[3], self.covs[3].shape[0] // 3)
            else:
                bap = bap_full[:, 0:self.num_bap]
            output_list.append(bap)
        return np.concatenate(output_list, axis=1)

    def _get_id_list(self, dir_in, file_id_list, id_list, file_ext):
        if id_list is None:
            id_list = [os.path.splitext(os.path.basename(f))[0] for f in sorted(glob.glob(os.path.join(dir_in, "*" + file_ext)))]
            file_id_list_name = "all"
        else:
            file_id_list_name = os.path.splitext(os.path.basename(file_id_list))[0]
        return id_list, file_id_list_name

    def _create_directories(self, dir_out):
        for load, d in ((self.load_sp, self.dir_coded_sps), (self.load_lf0, self.dir_lf0), (self.load_vuv, self.dir_vuv),
                        (self.load_bap, self.dir_bap)):
            if load:
                os.makedirs(os.path.join(dir_out, d), exist_ok=True)
