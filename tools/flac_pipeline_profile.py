"""Developer tool: where the time of the FLAC end-to-end leg goes: per-step wall time distribution and the time spent
waiting for the pack (oe_flac_wait), enqueueing copies + decode (to_device), waiting for the decode (drop_failed)."""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from openeat_b200 import _lib
from openeat_b200 import ingest as ING
from openeat_b200.dataset import PrefetchingCollator, audio_collate_func
from tools.flac_gpu_bench import encode, speechlike

dev = torch.device('cuda', 0)
lib = _lib.load()
lens, speeds = bench.workload(0)
d = '/dev/shm/oe_flac_prof'
os.makedirs(d, exist_ok=True)
rng = np.random.default_rng(0)
keys = ['u%d' % i for i in range(bench.BATCH)]
labels = [[1, 2, 3]] * bench.BATCH
batches = []
for bi in range(3):
    items = []
    for u in range(bench.BATCH):
        p = os.path.join(d, 'b%d_u%d.flac' % (bi, u))
        with open(p, 'wb') as f:
            f.write(encode(lib, speechlike(rng, int(lens[u]))))
        items.append((keys[u], p, labels[u], speeds[u]))
    batches.append(items)
mean = torch.linspace(8.0, 12.0, 80, device=dev)
istd = torch.linspace(0.4, 0.6, 80, device=dev)
stats = torch.zeros(161, dtype=torch.float64, device=dev)
collate = audio_collate_func(data_type='wav', feature_extraction_conf=bench.CONF, normalization=True, spec_aug=True,
                             spec_aug_conf=bench.AUG, global_cmvn=(mean, istd), cmvn_stats=stats)
acc = {}


def timed(cls, name):
    fn = getattr(cls, name)

    def wrap(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc.setdefault(name, []).append(time.perf_counter() - t0)
    setattr(cls, name, wrap)


for cls, name in ((ING.FlacGpuIngest, 'wait'), (ING.FlacGpuIngest, 'submit'), (ING.FlacGpuIngest, '_buffers'), (ING.FlacBatch, 'to_device'),
                  (ING.FlacBatch, 'drop_failed'), (audio_collate_func, 'collate_packed')):
    timed(cls, name)


def forever():
    i = 0
    while True:
        yield batches[i % 3]
        i += 1


for workers, threads in ((2, 16), (1, 16), (2, 8)):
    pipe = PrefetchingCollator(collate, ING.flac_gpu_batches(forever(), depth=3, workers=workers, threads=threads))
    pin_n = torch.empty(bench.BATCH, dtype=torch.int32).pin_memory()
    random.seed(1)
    for i in range(30):
        next(pipe)
    torch.cuda.synchronize()
    acc.clear()
    steps = []
    t0 = time.perf_counter()
    for i in range(200):
        _, out = next(pipe)
        pin_n.copy_(out['features_length'], non_blocking=True)
        t1 = time.perf_counter()
        steps.append(t1 - t0)
        t0 = t1
    torch.cuda.synchronize()
    s = np.array(steps) * 1e3
    print('workers=%d threads=%d: step ms median %.3f mean %.3f p90 %.3f max %.3f' % (workers, threads, np.median(s), s.mean(), np.percentile(s, 90), s.max()))
    for k, v in sorted(acc.items()):
        v = np.array(v) * 1e3
        print('   %-16s calls %4d  median %.3f  mean %.3f  max %.3f ms' % (k, len(v), np.median(v), v.mean(), v.max()))
    del pipe
