"""The sharded product path on real GPUs: WorldFeatLabelGen.gen_data under torch.distributed (NCCL), two ranks.  Needs two
devices (run with `gpurun --gpus 2`; skipped otherwise).  Checks that the all-reduced statistics equal the single-GPU ones,
that every utterance is written exactly once, and that a failure on ONE rank raises on BOTH (no hang in the collective)."""
import os
import socket
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_corpus(root, n, fs):
    from idiaptts_b200 import synthetic
    waves, f0s = synthetic.make_corpus(n, fs, seed=11, mean_dur=0.7, std_dur=0.3)
    os.makedirs(os.path.join(root, "wav"), exist_ok=True)
    ids = []
    for u, (w, f) in enumerate(zip(waves, f0s)):
        id_ = "utt%03d" % u
        with wave.open(os.path.join(root, "wav", id_ + ".wav"), "wb") as wf:
            wf.setnchannels(1)
            wf.setsampwidth(2)
            wf.setframerate(fs)
            wf.writeframes(w.numpy().tobytes())
        np.save(os.path.join(root, "wav", id_ + ".npy"), f)
        ids.append(id_)
    return ids


def _worker(rank, world, port, root, ids, break_rank):
    import torch.distributed as dist
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cache = {i: np.load(os.path.join(root, "wav", i + ".npy")) for i in ids}
    out = os.path.join(root, "out_sharded" if break_rank < 0 else "out_broken")
    if rank == break_rank:
        from idiaptts_b200 import distributed
        lens = []
        for i in ids:
            with wave.open(os.path.join(root, "wav", i + ".wav"), "rb") as w:
                lens.append(w.getnframes())
        victim = ids[int(distributed.shard_utterances(lens, world)[rank][0])]
        cache[victim] = cache[victim][:-3]        # wrong frame count on THIS rank's shard only
    gen = WorldFeatLabelGen(out, num_coded_sps=60, num_bap=2, f0_cache=cache)
    try:
        means, stds = gen.gen_data(os.path.join(root, "wav"), out, file_id_list="train.txt", id_list=ids)
        np.savez(os.path.join(root, "result_rank%d.npz" % rank), means=means, stds=stds)
    except Exception as e:  # noqa: BLE001
        with open(os.path.join(root, "error_rank%d.txt" % rank), "w") as f:
            f.write(type(e).__name__ + ": " + str(e))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_gen_data_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    fs = 22050
    root = str(tmp_path)
    ids = _make_corpus(root, 7, fs)
    cache = {i: np.load(os.path.join(root, "wav", i + ".npy")) for i in ids}
    single = WorldFeatLabelGen(os.path.join(root, "out_single"), num_coded_sps=60, num_bap=2, f0_cache=cache)
    m1, s1 = single.gen_data(os.path.join(root, "wav"), os.path.join(root, "out_single"), file_id_list="train.txt", id_list=ids)
    mp.spawn(_worker, args=(2, _free_port(), root, ids, -1), nprocs=2, join=True)
    r0, r1 = np.load(os.path.join(root, "result_rank0.npz")), np.load(os.path.join(root, "result_rank1.npz"))
    assert np.array_equal(r0["means"], r1["means"]) and np.array_equal(r0["stds"], r1["stds"])      # every rank holds the totals
    np.testing.assert_allclose(r0["means"], m1, rtol=1e-10, atol=1e-12)                            # fp64 sums, different order
    np.testing.assert_allclose(r0["stds"], s1, rtol=1e-8, atol=1e-10)
    for i in ids:                                                                                   # files: written once, identical
        for d, key in (("mcep60", "mcep"), ("lf0", "lf0"), ("vuv", "vuv"), ("bap", "bap")):
            a = np.load(os.path.join(root, "out_single", d, i + ".npz"))[key]
            b = np.load(os.path.join(root, "out_sharded", d, i + ".npz"))[key]
            assert np.array_equal(a, b), (i, d)
    assert os.path.exists(os.path.join(root, "out_sharded", "mcep60", "train-mean-std_dev.npz"))
    # a failure on rank 1 only: both ranks raise after the collective, nobody hangs
    mp.spawn(_worker, args=(2, _free_port(), root, ids, 1), nprocs=2, join=True)
    e0 = open(os.path.join(root, "error_rank0.txt")).read()
    e1 = open(os.path.join(root, "error_rank1.txt")).read()
    assert "other rank" in e0 and "cached F0" in e1
