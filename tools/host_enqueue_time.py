"""Developer tool: host time per prepared step (fe.run = oe_fbank_run: four launches).  TINY=1 shrinks the utterances so
that the GPU is faster than the host and the loop shows the host cost alone (with the full batch the pinned metadata
ring throttles the host to the GPU's pace)."""
import os, sys, time, random
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from openeat_b200 import planner
from openeat_b200.dataset import _plan_batch
from openeat_b200.frontend import default_frontend
from openeat_b200._lib import OE_WAV_I16
dev = torch.device('cuda', 0)
fe = default_frontend(80, 16000, dev)
lens, speeds = bench.workload(0)
if os.environ.get('TINY'):
    lens = np.full_like(lens, 1600)          # ~8 frames per utterance: the GPU step is short, what is left is the host
host_pool, offs = bench.synth_pool_host(lens, 0, 2)
dev_pool = [h.to(dev) for h in host_pool]
keys = ['u%d' % i for i in range(bench.BATCH)]; labels = [[1,2,3]] * bench.BATCH
mean = torch.linspace(8.0, 12.0, 80, device=dev); istd = torch.linspace(0.4, 0.6, 80, device=dev)
stats = torch.zeros(161, dtype=torch.float64, device=dev)
random.seed(1)
plan = _plan_batch(keys, labels, lens, [16000] * bench.BATCH, speeds, bench.CONF)
_, tm, fm = planner.plan_augment(plan.frames, 80, None, bench.AUG)
prep = fe.prepare(OE_WAV_I16, offs[plan.src], lens[plan.src], layout='padded', normalization=True, tmask=tm, fmask=fm,
                  cmvn=(mean, istd), cmvn_on_padding=True, stats=stats, speed_ratios=plan.stage2)
for i in range(20): fe.run(prep, dev_pool[i % 2])
torch.cuda.synchronize()
for n in (50, 150):
    t0 = time.perf_counter()
    for i in range(n): fe.run(prep, dev_pool[i % 2])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('n=%d host enqueue %.1f us per step, total %.1f us per step' % (n, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
