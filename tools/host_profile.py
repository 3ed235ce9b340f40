"""Developer tool: where does the HOST time of one bench step go (cProfile over the resident and e2e steps)."""
import cProfile
import os
import pstats
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from openeat_b200 import planner
from openeat_b200.dataset import _plan_batch, _run_plan, audio_collate_func
from openeat_b200.frontend import default_frontend

dev = torch.device('cuda', 0)
fe = default_frontend(80, 16000, dev)
lens, speeds = bench.workload(0)
host_pool, offs = bench.synth_pool_host(lens, 0, 2)
dev_pool = [h.to(dev) for h in host_pool]
keys = ['u%d' % i for i in range(bench.BATCH)]
labels = [[1, 2, 3]] * bench.BATCH
mean = torch.linspace(8.0, 12.0, 80, device=dev)
istd = torch.linspace(0.4, 0.6, 80, device=dev)
stats = torch.zeros(161, dtype=torch.float64, device=dev)
random.seed(1)
plan = _plan_batch(keys, labels, lens, [16000] * bench.BATCH, speeds, bench.CONF)
_, tm, fm = planner.plan_augment(plan.frames, 80, None, bench.AUG)
collate = audio_collate_func(data_type='wav', feature_extraction_conf=bench.CONF, normalization=True, spec_aug=True,
                             spec_aug_conf=bench.AUG, global_cmvn=(mean, istd), cmvn_stats=stats)


def resident(n):
    for i in range(n):
        _run_plan(plan, 80, dev_pool[i % 2], offs, lens, normalization=True, tmask=tm, fmask=fm, cmvn=(mean, istd),
                  cmvn_on_padding=True, stats=stats)


def e2e(n):
    for i in range(n):
        _, out = collate.collate_packed(host_pool[i % 2], offs, lens, keys, labels, speeds)
        out['features_length'].cpu(), stats.cpu()


from openeat_b200._lib import OE_WAV_I16   # noqa: E402
from openeat_b200.dataset import PrefetchingCollator   # noqa: E402

prep = fe.prepare(OE_WAV_I16, offs[plan.src], lens[plan.src], layout='padded', normalization=True, tmask=tm, fmask=fm,
                  cmvn=(mean, istd), cmvn_on_padding=True, stats=stats, speed_ratios=plan.stage2)


def prepared(n):
    for i in range(n):
        fe.run(prep, dev_pool[i % 2])


def batches():
    i = 0
    while True:
        yield (host_pool[i % 2], offs, lens, keys, labels, speeds)
        i += 1


pipe = PrefetchingCollator(collate, batches())
pin_n = torch.empty(bench.BATCH, dtype=torch.int32).pin_memory()
pin_s = torch.empty(161, dtype=torch.float64).pin_memory()


def pipelined(n):
    for i in range(n):
        _, out = next(pipe)
        pin_n.copy_(out['features_length'], non_blocking=True)
        pin_s.copy_(stats, non_blocking=True)


import tempfile   # noqa: E402
import wave       # noqa: E402
from openeat_b200.ingest import ingest_batches   # noqa: E402

wav_dir = tempfile.mkdtemp(dir='/dev/shm')
file_batches = []
for bi in range(2):
    pcm = host_pool[bi].numpy()
    items = []
    for u in range(bench.BATCH):
        path = os.path.join(wav_dir, 'b%d_u%d.wav' % (bi, u))
        with wave.open(path, 'wb') as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(16000)
            w.writeframes(pcm[offs[u]:offs[u] + lens[u]].tobytes())
        items.append((keys[u], path, labels[u], speeds[u]))
    file_batches.append(items)


def file_items():
    i = 0
    while True:
        yield file_batches[i % 2]
        i += 1


pipe_files = PrefetchingCollator(collate, ingest_batches(file_items(), depth=3))


def from_files(n):
    for i in range(n):
        _, out = next(pipe_files)
        pin_n.copy_(out['features_length'], non_blocking=True)
        pin_s.copy_(stats, non_blocking=True)


for name, fn in (('from wav files', from_files), ('pipelined e2e', pipelined)):
    fn(5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(50)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print('%s: host enqueue %.3f ms/step, with GPU drain %.3f ms/step' % (name, t_host / 50 * 1e3, t_all / 50 * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    fn(50)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
