"""Host logic of the multi-GPU path: utterances are independent, so ranks shard them with no data-path
collective; only the CMVN statistics are all-reduced (openeat_b200.cmvn)."""
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
