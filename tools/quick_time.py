"""Developer timing of the fbank kernel alone (not the bench): batch 256 of 2-10 s int16 utterances."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from openeat_b200.frontend import Frontend, aligned_offsets

peak = 6445.3
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
fe = Frontend()
rng = np.random.default_rng(1002)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lens = np.round(rng.uniform(2, 10, B) * 16000).astype(np.int32)
offs, total = aligned_offsets(lens)
pool = [torch.randint(-3000, 3000, (total,), dtype=torch.int16, device='cuda') for _ in range(6)]
frames = fe.num_frames_array(lens)
out = torch.empty((int(frames.sum()), 80), device='cuda')
alg_bytes = 2 * int(lens.sum()) + 320 * int(frames.sum())
for mode, kw in [('raw', {}), ('norm', dict(normalization=True))]:
    for dtype in ('i16', 'f32'):
        ps = pool if dtype == 'i16' else [p.float() for p in pool[:3]]
        ab = alg_bytes if dtype == 'i16' else alg_bytes + 2 * int(lens.sum())
        for i in range(3):
            fe.fbank(ps[i % len(ps)], offs, lens, layout='ragged', out=out, **kw)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n = 20
        ev[0].record()
        for i in range(n):
            fe.fbank(ps[i % len(ps)], offs, lens, layout='ragged', out=out, **kw)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n
        print('%s %s: %.3f ms/batch  %.3e frames/s  %.3e audio-s/s  %.1f GB/s alg = %.1f%% of %.0f GB/s' % (
            mode, dtype, ms, frames.sum() / ms * 1e3, lens.sum() / 16000 / ms * 1e3, ab / ms / 1e6,
            100 * ab / ms / 1e6 / peak, peak))
