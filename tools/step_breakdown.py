"""Developer timing on one box: the benchmark step's kernels, bracketed by the library's own CUDA events
(oe_frontend_set_kernel_timing: fbank kernel alone, and the whole launch sequence) -> time left for the rest."""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from openeat_b200 import planner  # noqa: E402
from openeat_b200.dataset import _plan_batch, _run_plan  # noqa: E402
from openeat_b200.frontend import default_frontend  # noqa: E402

dev = torch.device('cuda', 0)
fe = default_frontend(80, 16000, dev)
lens, speeds = bench.workload(0)
host_pool, offs = bench.synth_pool_host(lens, 0, bench.POOL)
dev_pool = [h.to(dev) for h in host_pool]
keys = ['utt%d' % i for i in range(bench.BATCH)]
labels = [[1, 2, 3]] * bench.BATCH
mean = torch.linspace(8.0, 12.0, 80, device=dev)
istd = torch.linspace(0.4, 0.6, 80, device=dev)
stats = torch.zeros(161, dtype=torch.float64, device=dev)
random.seed(4242)
plans = []
for _ in range(bench.POOL):
    plan = _plan_batch(keys, labels, lens, [16000] * bench.BATCH, speeds, bench.CONF)
    _, tm, fm = planner.plan_augment(plan.frames, 80, None, bench.AUG)
    plans.append((plan, tm, fm))


def step(i):
    plan, tm, fm = plans[i % bench.POOL]
    _run_plan(plan, 80, dev_pool[i % bench.POOL], offs, lens, normalization=True, tmask=tm, fmask=fm,
              cmvn=(mean, istd), cmvn_on_padding=True, stats=stats)


for i in range(5):
    step(i)
torch.cuda.synchronize()
fe.set_kernel_timing(True)
fb, st = [], []
for i in range(30):
    step(i)
    fb.append(fe.fbank_kernel_ms())
    st.append(fe.step_ms())
fe.set_kernel_timing(False)
print('fbank kernel %.1f us   whole launch sequence %.1f us   -> descriptors + completion %.1f us (events end the PDL overlap)'
      % (1e3 * np.median(fb), 1e3 * np.median(st), 1e3 * (np.median(st) - np.median(fb))))
