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


def gpu_numa_node(index):
    """NUMA node the PCIe root complex of GPU ``index`` hangs off, from sysfs (``/sys/bus/pci/devices/<bdf>/numa_node``);
    None when the platform does not say (single-node hosts and most virtual machines report -1)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % bdf) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def _node_cpus(node):
    cpus = set()
    with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
        for part in f.read().strip().split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(index):
    """Places the calling process next to GPU ``index``: CPU affinity AND memory policy (``set_mempolicy(MPOL_PREFERRED)``,
    so that the pinned staging rings allocated afterwards live in the memory of the GPU's own PCIe root complex and
    host-to-device copies do not cross the socket interconnect).  The node comes from sysfs (``gpu_numa_node``); NVML's
    CPU affinity is the fallback (it names CPUs only, and some hosts report one set for every GPU).  One process per
    GPU (the reference's DDP launch, train_ddp.py) should call it before allocating pinned memory.  Best effort:
    returns a short description of what was bound, or None."""
    node = gpu_numa_node(index)
    if node is not None:
        try:
            import ctypes
            cpus = _node_cpus(node) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))     # x86-64 set_mempolicy, MPOL_PREFERRED
            return 'numa node %d (sysfs): %d cpus%s' % (node, len(cpus), ', memory preferred' if rc == 0 else '')
        except Exception:
            pass
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and cpus != os.sched_getaffinity(0):
            os.sched_setaffinity(0, cpus)
            return '%d cpus (NVML affinity)' % len(cpus)
    except Exception:
        pass
    return None
