"""Multi-GPU plumbing of the WORLD feature path: utterances are independent units, so they are dealt to ranks with no
data-path collective; the ONE exchange step is the sum of the corpus normalisation statistics, i.e. the semantics of
MeanStdDevExtractor.combine_stats (idiaptts/misc/normalisation/MeanStdDevExtractor.py:163-204) as one all-reduce
(NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_utterances(lengths, world_size):
    """Greedy longest-first assignment of utterances to ranks, balancing the total length (frames / samples).
    Returns a list (one entry per rank) of index arrays; deterministic, every utterance assigned exactly once."""
    lengths = np.asarray(lengths, np.int64)
    order = np.argsort(-lengths, kind="stable")
    loads = np.zeros(world_size, np.int64)
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(loads))
        shards[r].append(int(i))
        loads[r] += lengths[i]
    return [np.array(sorted(s), np.int64) for s in shards]


def allreduce_stats(buf):
    """In-place sum over ranks of the packed fp64 statistics buffer [N, sum x (D), sum x^2 (D), ...]."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf
