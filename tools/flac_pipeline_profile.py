"""Developer tool: where the time of the FLAC end-to-end leg goes (host profile of the pipelined step, pack timings)."""
import cProfile
import ctypes
import os
import pstats
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from openeat_b200 import _lib
from openeat_b200.dataset import PrefetchingCollator, audio_collate_func
from openeat_b200.ingest import FlacGpuIngest, flac_gpu_batches
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


def forever():
    i = 0
    while True:
        yield batches[i % 3]
        i += 1


for workers, threads in ((1, 16), (2, 16), (2, 8), (4, 16)):
    g = FlacGpuIngest(threads=threads // workers, ring=3)
    for _ in range(3):
        g.pack([x[1] for x in batches[0]], report=False)
    t0 = time.perf_counter()
    for _ in range(20):
        g.pack([x[1] for x in batches[0]], report=False)
    print('pack alone, %2d threads: %.2f ms' % (threads // workers, (time.perf_counter() - t0) / 20 * 1e3))
    del g
    pipe = PrefetchingCollator(collate, flac_gpu_batches(forever(), depth=3, workers=workers, threads=threads))
    pin_n = torch.empty(bench.BATCH, dtype=torch.int32).pin_memory()

    def run(n):
        for i in range(n):
            _, out = next(pipe)
            pin_n.copy_(out['features_length'], non_blocking=True)

    random.seed(1)
    run(20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(100)
    torch.cuda.synchronize()
    print('pipeline, workers=%d threads=%d: %.3f ms per step' % (workers, threads, (time.perf_counter() - t0) / 100 * 1e3))
    if workers == 2 and threads == 16:
        pr = cProfile.Profile()
        pr.enable()
        run(100)
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr).sort_stats('tottime').print_stats(18)
