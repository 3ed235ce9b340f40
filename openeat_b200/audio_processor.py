"""Drop-in for ``openeat/dataset/audio_processor.py`` (same names, arguments, RNG behaviour).

``_speed_generator`` is host logic and keeps the reference's exact ``random`` call pattern
(including its quirk of always returning ``speeds[0]`` when the step is non-zero, SURVEY
appendix A.1).  ``_speed_perturb`` runs the polyphase sinc resampler on the GPU
(``oe_resample``): libsox ``speed s`` + ``rate sr`` == resample from ``int(s*sr)`` to ``sr``.
"""
import random

import numpy as np
import torch

from .frontend import aligned_offsets, default_frontend, speed_ratio


def _speed_generator(speeds):
    """openeat/dataset/audio_processor.py:5-18."""
    if speeds is None:
        speeds = [0.9, 1.1, 0.1]
    speeds = [float(s) for s in speeds]
    if len(speeds) > 1:
        assert speeds[1] > speeds[0], 'speeds is wrong !'
        if speeds[2] != 0:
            speed = random.randrange(int(speeds[0] / speeds[2]), int(speeds[0] / speeds[2]) + 1)
            speed *= speeds[2]
        else:
            speed = speeds[0] + random.random() * (speeds[1] - speeds[0])
    else:
        speed = speeds[0]
    return speed


def _speed_perturb(waveform, sample_rate, speed=None):
    """openeat/dataset/audio_processor.py:19-35.  ``waveform`` is a (1, N) float tensor on the int16
    scale; returns the input object itself for ``speed == 1.0`` and a new (1, ~N/speed) tensor on the
    input's device otherwise."""
    if speed == 1.0:
        return waveform
    fe = default_frontend(sample_rate=16000)
    src_device = waveform.device
    x = waveform.reshape(-1).to(device=fe.device, dtype=torch.float32).contiguous()
    n = x.shape[0]
    offs, total = aligned_offsets([n])
    if total != n:
        x = torch.nn.functional.pad(x, (0, total - n))
    out, ooffs, olens = fe.resample(x, offs, np.array([n], np.int32), [speed_ratio(speed, sample_rate)])
    return out[ooffs[0]:ooffs[0] + int(olens[0])].unsqueeze(0).to(src_device)
