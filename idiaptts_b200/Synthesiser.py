"""Synthesiser.run_world_synth of the reference (idiaptts/src/Synthesiser.py:39-80) as ONE batched GPU call: all
utterances of `synth_output` are decoded (convert_to_world_features -> decode_sp -> world_features_to_raw) together by
pipeline.WorldSynthesizer and written as PCM16 wav files named like the reference's
(<id><synth_file_suffix>_<num_coded_sps><sp_type>_WORLD.wav)."""
import logging
import os
import wave

import numpy as np
import torch

from . import corpus_io, pipeline
from .WorldFeatLabelGen import WorldFeatLabelGen


class Synthesiser(object):
    SYNTH_SUB_DIR = "synth"

    @staticmethod
    def synth_batch(synth_output, hparams, has_deltas=False):
        """{id: [T, D]} -> {id: waveform float32}; the compute half of run_world_synth."""
        ids, y, out_off = Synthesiser._synth_packed(synth_output, hparams, has_deltas)
        return {ids[u]: y[out_off[u]:out_off[u + 1]] for u in range(len(ids))}

    @staticmethod
    def _synth_packed(synth_output, hparams, has_deltas=False):
        """-> (ids, packed float32 waveforms on the host, sample offsets [U + 1])."""
        if not torch.cuda.is_available():
            raise RuntimeError("idiaptts_b200 needs a CUDA device; there is no CPU fallback")
        sp_type = getattr(hparams, "sp_type", "mcep")
        if sp_type not in ("mcep", "mgc"):
            raise NotImplementedError("sp_type '{}': only 'mcep' and 'mgc' are on the accelerated path".format(sp_type))
        dev = torch.device("cuda", torch.cuda.current_device())
        D, nb, fs = hparams.num_coded_sps, hparams.num_bap, hparams.synth_fs
        rows, lens, ids = [], [], []
        for id_name, output in synth_output.items():
            coded_sp, lf0, vuv, bap = WorldFeatLabelGen.convert_to_world_features(np.asarray(output), contains_deltas=has_deltas,
                                                                                  num_coded_sps=D, num_bap=nb)
            rows.append(np.concatenate((coded_sp, lf0[:, None], vuv[:, None], bap.reshape(len(lf0), -1)), axis=1).astype(np.float32))
            lens.append(len(lf0))
            ids.append(id_name)
        if not rows:
            return [], np.zeros(0, np.float32), np.zeros(1, np.int64)
        syn = pipeline.WorldSynthesizer(fs, D, getattr(hparams, "mgc_alpha", None),
                                        f0_silence_threshold=getattr(hparams, "f0_silence_threshold", WorldFeatLabelGen.f0_silence_threshold),
                                        lf0_zero=getattr(hparams, "lf0_zero", WorldFeatLabelGen.lf0_zero), device=dev,
                                        sp_type=sp_type, mgc_gamma=getattr(hparams, "mgc_gamma", None) or -1.0 / 3.0,
                                        post_filtering=getattr(hparams, "do_post_filtering", False))
        feats = torch.from_numpy(np.concatenate(rows)).to(dev)
        frame_off = torch.from_numpy(np.concatenate(([0], np.cumsum(lens))).astype(np.int64)).to(dev)
        y, out_off, status = syn.synthesize(feats, frame_off, preemphasis=getattr(hparams, "preemphasis", 0.0))
        y = y.cpu().numpy()
        from . import ops
        ops.raise_for_status(status, "run_world_synth")
        return ids, y, np.asarray(out_off, np.int64)

    @staticmethod
    def run_world_synth(synth_output, hparams, epoch=None, step=None, use_model_name=True, has_deltas=False):
        save_dir = Synthesiser._get_synth_dir(hparams, use_model_name, epoch=epoch, step=step)
        if getattr(hparams, "synth_ext", "wav").lower() != "wav":
            raise NotImplementedError("only wav output is supported (pydub re-encoding is outside the path)")
        ids, y, out_off = Synthesiser._synth_packed(synth_output, hparams, has_deltas)
        paths = []
        for id_name in ids:
            logging.info("Synthesise {} with the WORLD vocoder.".format(id_name))
            file_name = (os.path.basename(id_name) + getattr(hparams, "synth_file_suffix", "") + "_" + str(hparams.num_coded_sps)
                         + hparams.sp_type + "_WORLD")
            paths.append(os.path.join(save_dir, file_name + ".wav"))
        if y.dtype == np.float32:
            # all files of the batch in one native call (a pool of host threads; same bytes as write_wav)
            corpus_io.write_wavs_pcm16(paths, np.ascontiguousarray(y), out_off, hparams.synth_fs)
        else:  # float64 samples (de-emphasis): quantised from the doubles, file by file
            for u, path in enumerate(paths):
                Synthesiser.write_wav(path, y[out_off[u]:out_off[u + 1]], hparams.synth_fs)

    @staticmethod
    def write_wav(path, waveform, fs):
        pcm = np.clip(np.round(np.asarray(waveform, np.float64) * 32767.0), -32768, 32767).astype(np.int16)
        with wave.open(path, "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(int(fs))
            w.writeframes(pcm.tobytes())

    @staticmethod
    def _get_synth_dir(hparams, use_model_name=True, epoch=None, step=None):
        def has(name):
            return getattr(hparams, name, None) is not None
        if has("synth_dir"):
            save_dir = hparams.synth_dir
        else:
            parts = [hparams.out_dir] if has("out_dir") else [os.path.curdir]
            if use_model_name and has("model_name"):
                parts.append(hparams.model_name)
            parts.append(Synthesiser.SYNTH_SUB_DIR)
            if epoch is not None:
                parts.append("e" + str(epoch))
            elif step is not None:
                parts.append("s" + str(step))
            save_dir = os.path.join(*parts)
        os.makedirs(save_dir, exist_ok=True)
        return save_dir
