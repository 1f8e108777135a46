"""Objective metrics on the device against the numpy restatement of src/Metrics.py."""
import numpy as np
import pytest
import torch

from oracle import glue_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nap", [1, 2, 5])
def test_batch_metrics_match_oracle(nap):
    from idiaptts_b200.Metrics import Metrics
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(nap)
    D, lens = 60, [37, 120, 5]
    rows_o, rows_x, fu = [], [], []
    ref = []
    for u, T in enumerate(lens):
        sp = rng.standard_normal((T, D)).astype(np.float32)
        lf0 = rng.uniform(4.0, 6.0, (T, 1)).astype(np.float32)
        vuv = (rng.uniform(0, 1, (T, 1)) > 0.3).astype(np.float32)
        vuv[0] = 1.0
        bap = -rng.uniform(0, 20, (T, nap)).astype(np.float32)
        sp2 = (sp + 0.1 * rng.standard_normal(sp.shape)).astype(np.float32)
        lf02 = (lf0 * (1 + 0.15 * rng.standard_normal(lf0.shape))).astype(np.float32)
        vuv2 = np.where(rng.uniform(0, 1, vuv.shape) > 0.9, 1 - vuv, vuv).astype(np.float32)
        vuv2[0] = 1.0
        bap2 = (bap + rng.standard_normal(bap.shape)).astype(np.float32)
        rows_o.append(np.concatenate([sp, lf0, vuv, bap], 1))
        rows_x.append(np.concatenate([sp2, lf02, vuv2, bap2], 1))
        fu.append(np.full(T, u, np.int32))
        ref.append(glue_np.world_metrics(sp, lf0[:, 0], vuv[:, 0], bap if nap > 1 else bap[:, 0], sp2, lf02[:, 0], vuv2[:, 0],
                                         bap2 if nap > 1 else bap2[:, 0]))
    res = Metrics.batch(torch.from_numpy(np.concatenate(rows_o)).to(dev), torch.from_numpy(np.concatenate(rows_x)).to(dev),
                        torch.from_numpy(np.concatenate(fu)).to(dev), len(lens), D, nap)
    for u in range(len(lens)):
        for k, v in ref[u].items():
            assert abs(res[k][u] - v) <= 1e-6 * max(1.0, abs(v)), (k, u, res[k][u], v)
    # the reference-facing single-utterance call
    T = lens[1]
    o, x = rows_o[1], rows_x[1]
    got = dict(Metrics.get_metrics([Metrics.MCD, Metrics.VDE, Metrics.F0_RMSE], org_coded_sp=o[:, :D], org_lf0=o[:, D], org_vuv=o[:, D + 1],
                                   org_bap=o[:, D + 2:], output_coded_sp=x[:, :D], output_lf0=x[:, D], output_vuv=x[:, D + 1],
                                   output_bap=x[:, D + 2:]))
    for k in got:
        assert abs(got[k] - ref[1][k]) <= 1e-6 * max(1.0, abs(ref[1][k]))
