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


def bind_host_to_gpu(device_index):
    """One process per GPU moves GBs per second between pinned host memory and its GPU (wav in, features / waveforms out): on a
    two-socket node the process and its page-locked buffers should live on the socket the GPU hangs off, or every copy crosses the
    inter-socket link.  Restricts the calling process to the CPUs NVML names as ideal for the device, intersected with the CPUs it is
    allowed to use; memory touched afterwards is then allocated on that node (first touch).  Returns the CPU set it bound to, or
    None when nothing was changed (single node, NVML unavailable, B2W_NUMA_BIND=0).  Call it before allocating pinned memory."""
    import os
    if os.environ.get("B2W_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        handle = None
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids):
                tok = ids[device_index]
                handle = pynvml.nvmlDeviceGetHandleByIndex(int(tok)) if tok.isdigit() else pynvml.nvmlDeviceGetHandleByUUID(tok.encode())
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        allowed = os.sched_getaffinity(0)
        words = (max(allowed) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        ideal = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus = ideal & allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 -- an optimisation only: never fatal
        return None
