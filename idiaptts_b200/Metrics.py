"""Objective metrics of the reference's Metrics class (idiaptts/src/Metrics.py:84-164) computed on the GPU for whole batches
(`Metrics.batch`) and, with the reference's own signatures, for single utterances (numpy in, float out).  No CPU fallback."""
import math

import numpy as np
import torch

from . import ops

_MCD_SCALE = 10.0 / math.log(10.0) * math.sqrt(2.0)


class Metrics(object):
    MCD = "MCD"
    F0_RMSE = "F0 RMSE"
    GPE = "GPE"
    FFE = "FFE"
    VDE = "VDE"
    BAP_distortion = "BAP distortion"

    @staticmethod
    def from_sums(acc, num_bap):
        """acc [U, 8] (ops.world_metrics) -> dict of per-utterance numpy arrays."""
        a = np.asarray(acc, np.float64)
        T = a[:, 7]
        gpe_num = a[:, 3]
        with np.errstate(divide="ignore", invalid="ignore"):
            res = {Metrics.MCD: _MCD_SCALE * a[:, 0] / T,
                   Metrics.F0_RMSE: np.sqrt(a[:, 1] / a[:, 2]),
                   Metrics.GPE: gpe_num / a[:, 4],
                   Metrics.VDE: a[:, 5] / T,
                   Metrics.FFE: gpe_num / T + a[:, 5] / T,
                   Metrics.BAP_distortion: (_MCD_SCALE * a[:, 6] / T) if num_bap > 1 else np.sqrt(a[:, 6] / T) * _MCD_SCALE}
        return res

    @staticmethod
    def batch(org, out, frame_utt, num_utts, num_coded_sps, num_bap):
        """Device tensors [F, D + 2 + nap] float32 (rows [coded_sp | lf0 | vuv | bap]) -> dict of per-utterance metrics."""
        acc = ops.world_metrics(org, out, frame_utt, num_utts, num_coded_sps, num_bap)
        return Metrics.from_sums(acc.cpu().numpy(), num_bap)

    @staticmethod
    def _one(org_coded_sp, org_lf0, org_vuv, org_bap, output_coded_sp, output_lf0, output_vuv, output_bap):
        if not torch.cuda.is_available():
            raise RuntimeError("idiaptts_b200 needs a CUDA device; there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        T = len(output_coded_sp)  # the reference trims the original to the output's length

        def pack(c, l, v, b):
            return np.concatenate([np.asarray(c, np.float32)[:T], np.asarray(l, np.float32).reshape(-1, 1)[:T],
                                   np.asarray(v, np.float32).reshape(-1, 1)[:T], np.asarray(b, np.float32).reshape(len(b), -1)[:T]], axis=1)
        o, x = pack(org_coded_sp, org_lf0, org_vuv, org_bap), pack(output_coded_sp, output_lf0, output_vuv, output_bap)
        D, nap = np.shape(output_coded_sp)[1], o.shape[1] - np.shape(output_coded_sp)[1] - 2
        fu = torch.zeros(T, dtype=torch.int32, device=dev)
        res = Metrics.batch(torch.from_numpy(o).to(dev), torch.from_numpy(x).to(dev), fu, 1, D, nap)
        return {k: float(v[0]) for k, v in res.items()}

    @staticmethod
    def get_metrics(metric_names, org_coded_sp=None, org_lf0=None, org_vuv=None, org_bap=None, output_coded_sp=None,
                    output_lf0=None, output_vuv=None, output_bap=None):
        """Same call and return convention as Metrics.get_metrics (:42-82) for the WORLD metrics."""
        vals = Metrics._one(org_coded_sp, org_lf0, org_vuv, org_bap, output_coded_sp, output_lf0, output_vuv, output_bap)
        res = []
        for name in metric_names:
            if name not in vals:
                raise NotImplementedError("Unknown metric {}.".format(name))
            res.append((name, vals[name]))
        return res
