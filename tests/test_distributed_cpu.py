"""The N > 1 path on CPU: world_size-2 gloo.  Each rank takes its shard of the utterances (distributed.shard_utterances),
accumulates the packed statistics buffer and all-reduces it; the result must equal the single-process statistics and the
reference's combine_stats semantics."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from idiaptts_b200 import distributed, pipeline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IDS = ["LJ001-%04d" % i for i in range(1, 10)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ljspeech_world_golden.npz"))
    feats = [np.concatenate((g[i + "/cmp"][:, :20], g[i + "/cmp"][:, 60:61], g[i + "/cmp"][:, 63:65]), axis=1) for i in IDS]
    lens = [len(f) for f in feats]
    mine = distributed.shard_utterances(lens, world)[rank]
    dim = feats[0].shape[1]
    buf = torch.zeros(1 + 2 * dim, dtype=torch.float64)
    for u in mine:
        x = torch.from_numpy(feats[u].astype(np.float64))
        buf[0] += len(x)
        buf[1:1 + dim] += x.sum(0)
        buf[1 + dim:] += (x ** 2).sum(0)
    r, w = distributed.world_info()
    assert (r, w) == (rank, world)
    distributed.allreduce_stats(buf)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), buf.numpy())
    dist.destroy_process_group()


def test_two_rank_statistics_allreduce(tmp_path, golden):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)
    allf = np.concatenate([np.concatenate((golden[i + "/cmp"][:, :20], golden[i + "/cmp"][:, 60:61], golden[i + "/cmp"][:, 63:65]), axis=1)
                           for i in IDS]).astype(np.float64)
    dim = allf.shape[1]
    assert a[0] == 11579
    np.testing.assert_allclose(a[1:1 + dim], allf.sum(0), rtol=1e-12)
    np.testing.assert_allclose(a[1 + dim:], (allf ** 2).sum(0), rtol=1e-12)
    mean, std = pipeline.mean_std_from_sums(a[1:], a[0], dim)
    ref = golden["stats/mcep20/mean-std_dev/data"]
    np.testing.assert_allclose(mean[:20], ref[0], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(std[:20], ref[1], rtol=2e-5, atol=2e-6)
