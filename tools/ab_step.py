"""Developer A/B timing on ONE box: the bench's device-resident step and the fbank call alone, for the library
named by OE_LIB_PATH (default: the in-tree build).  Prints medians of many short repeats.

    for lib in a.so b.so; do OE_LIB_PATH=$lib python tools/ab_step.py; done
"""
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
if os.environ.get('AB_NO_RS'):
    speeds = [1.0] * len(speeds)
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
frames = fe.num_frames_array(lens)
out = torch.empty((int(frames.sum()), 80), device=dev)


def step(i):
    plan, tm, fm = plans[i % bench.POOL]
    _run_plan(plan, 80, dev_pool[i % bench.POOL], offs, lens, normalization=True, tmask=tm, fmask=fm,
              cmvn=(mean, istd), cmvn_on_padding=True, stats=stats)


def fbank_only(i):
    fe.fbank(dev_pool[i % bench.POOL], offs, lens, layout='ragged', out=out)


def med(fn, n=20, reps=int(os.environ.get('AB_REPS', '40'))):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    r = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        r.append(a.elapsed_time(b) / n * 1e3)
    return float(np.median(r)), float(np.min(r))


def step_nonorm(i):          # single pass: fused speed perturb + masks + CMVN straight into the padded tensor
    plan, tm, fm = plans[i % bench.POOL]
    _run_plan(plan, 80, dev_pool[i % bench.POOL], offs, lens, normalization=False, tmask=tm, fmask=fm,
              cmvn=(mean, istd), cmvn_on_padding=True, stats=None)


def step_nostats(i):
    plan, tm, fm = plans[i % bench.POOL]
    _run_plan(plan, 80, dev_pool[i % bench.POOL], offs, lens, normalization=True, tmask=tm, fmask=fm,
              cmvn=(mean, istd), cmvn_on_padding=True, stats=None)


if os.environ.get('AB_PARTS'):
    print('single-pass (rs + masks + cmvn, padded out) %.1f us | two-phase without global stats %.1f us' % (
        med(step_nonorm)[0], med(step_nostats)[0]))
s_med, s_min = med(step)
f_med, f_min = med(fbank_only)
print('%-40s step %.1f us (min %.1f)   fbank-only %.1f us (min %.1f)' % (
    os.path.basename(os.environ.get('OE_LIB_PATH', 'in-tree')), s_med, s_min, f_med, f_min))
