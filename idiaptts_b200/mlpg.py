"""Drop-in for idiaptts.misc.mlpg.MLPG (the reference's bandmat-based maximum-probability parameter generation,
idiaptts/misc/mlpg.py): same class, same `generation(features, covariance, feature_dim)` signature and result (float64
[frames, feature_dim]), computed by the CUDA pentadiagonal solver of libb200world (`b2w_mlpg`, all dimensions of the call in one
launch; `generation_batch` does all utterances of a ragged batch at once).  No CPU fallback."""
import numpy as np
import torch

from . import ops


class MLPG(object):
    def generation(self, features, covariance, feature_dim):
        if not torch.cuda.is_available():
            raise RuntimeError("idiaptts_b200 needs a CUDA device; there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        features = np.ascontiguousarray(features)
        if features.dtype not in (np.float32, np.float64):
            features = features.astype(np.float64)
        frames = features.shape[0]
        if features.shape[1] < 3 * feature_dim:
            raise ValueError("features need 3 * feature_dim columns [static | delta | delta-delta]")
        var3 = np.ascontiguousarray(np.diag(np.asarray(covariance, np.float64))[:3 * feature_dim])
        off = torch.tensor([0, frames], dtype=torch.int64, device=dev)
        out = ops.mlpg(torch.from_numpy(features).to(dev), torch.from_numpy(var3).to(dev), off, feature_dim)
        return out.cpu().numpy()

    @staticmethod
    def generation_batch(features, var3, frame_off, feature_dim):
        """Device tensors in, device tensor out: features [F, 3 D], var3 [3 D] f64, frame_off int64 [U + 1]."""
        return ops.mlpg(features, var3, frame_off, feature_dim)
