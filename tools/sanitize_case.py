"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel, few tiles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from openeat_b200.frontend import Frontend, pack_waveforms
from oracle import signals

fe = Frontend()
lens = [16000, 9000, 300, 5200, 12345, 400]
waves = [signals.make('speech', n, 900 + i) for i, n in enumerate(lens)]
buf, offs, ln = pack_waveforms(waves)
dev = buf.cuda()
ratios = np.array([[9, 10], [0, 0], [0, 0], [11, 10], [9, 10], [0, 0]])
stats = torch.zeros(161, dtype=torch.float64, device='cuda')
mean = torch.linspace(8.0, 12.0, 80, device='cuda')
istd = torch.linspace(0.4, 0.6, 80, device='cuda')
a, fr = fe.fbank(dev, offs, ln, layout='padded')                                              # single pass
b, _ = fe.fbank(dev, offs, ln, layout='padded', tmask=np.array([[[3, 9]]] * 6, np.int32), cmvn=(mean, istd))
c, fr2 = fe.fbank(dev, offs, ln, layout='padded', normalization=True, speed_ratios=ratios, stats=stats,
                  fmask=np.array([[[10, 14]]] * 6, np.int32), cmvn=(mean, istd), cmvn_on_padding=True)
r, ro, rl = fe.resample(dev, offs, ln, ratios)
d, _ = fe.fbank(r, ro, rl, layout='ragged')                                                   # fp32 input path
# spec_sub (scratch + out-of-place finalize), feature dither, the on-the-fly and Kaiser resamplers, 23 mel bins (gen-1 kernel)
maps = [np.arange(int(t), dtype=np.int32)[::-1].copy() for t in fr]
e, _ = fe.fbank(dev, offs, ln, layout='padded', normalization=True, frame_maps=maps, feature_dither=0.3, dither_seed=5)
r2, _, _ = fe.resample(dev, offs, ln, [(441, 160)] * 6)
r3, _, _ = fe.resample(dev, offs, ln, [(9, 10)] * 6, kind='kaiser')
fe23 = Frontend(mel_bins=23)
f23, _ = fe23.fbank(dev, offs, ln, layout='padded', normalization=True, stats=torch.zeros(47, dtype=torch.float64, device='cuda'))
up = fe.upload_small(np.arange(37, dtype=np.int32))
torch.cuda.synchronize()
assert up.tolist() == list(range(37))
print('ok', fr.tolist(), fr2.tolist(), float(stats[160]), bool(torch.isfinite(a).all()), bool(torch.isfinite(d).all()))
