"""Drop-in for ``openeat/dataset/feature_processor.py``: numpy (T, F) in, NEW numpy array out, random
indices drawn from Python's global ``random`` in the reference's exact order -- the index work and
the normalisation run on the GPU through the features-in mode of ``oe_fbank_batch``.

``plan_*`` expose the host-side index draws so whole batches can be planned first and executed in
one fused launch (what ``openeat_b200.dataset.audio_collate_func`` does).
"""
import random

import numpy as np
import torch

from .frontend import default_frontend

SPEC_MASK = 0  # feature_processor.py:4


def plan_spec_augmentation(num_frames, num_freq, num_t_mask=2, num_f_mask=2, max_t=50, max_f=10):
    """The random draws of feature_processor.py:31-41 -> ([t ranges], [f ranges]), half open, clipped."""
    t_masks, f_masks = [], []
    for _ in range(num_t_mask):
        start = random.randint(0, num_frames - 1)
        length = random.randint(1, max_t)
        t_masks.append((start, min(num_frames, start + length)))
    for _ in range(num_f_mask):
        start = random.randint(0, num_freq - 1)
        length = random.randint(1, max_f)
        f_masks.append((start, min(num_freq, start + length)))
    return t_masks, f_masks


def plan_spec_substitute(num_frames, max_t=20, num_t_sub=3):
    """The random draws of feature_processor.py:57-63, composed into one frame-index map: the copies
    move whole rows, so their sequence is a composition of index maps (y[t] = x[idx[t]])."""
    idx = np.arange(num_frames, dtype=np.int32)
    for _ in range(num_t_sub):
        start = random.randint(0, num_frames - 1)
        length = random.randint(1, max_t)
        end = min(num_frames, start + length)
        pos = random.randint(0, start)
        idx[start:end] = idx[start - pos:end - pos].copy()
    return idx


def _run(x, **kw):
    x = np.ascontiguousarray(x, dtype=np.float32)
    fe = default_frontend(mel_bins=x.shape[1])
    out, _ = fe.fbank(torch.from_numpy(x).to(fe.device), np.array([0], np.int64),
                      np.array([x.shape[0]], np.int32), layout='ragged', features_in=True, **kw)
    return out.cpu().numpy()


def _normalization(feature):
    """feature_processor.py:5-8."""
    return _run(feature, normalization=True)


def _spec_augmentation(x, num_t_mask=2, num_f_mask=2, max_t=50, max_f=10):
    """feature_processor.py:10-42."""
    t, f = plan_spec_augmentation(x.shape[0], x.shape[1], num_t_mask, num_f_mask, max_t, max_f)
    return _run(x, tmask=np.array([t], np.int32) if t else None, fmask=np.array([f], np.int32) if f else None)


def _spec_substitute(x, max_t=20, num_t_sub=3):
    """feature_processor.py:44-64."""
    return _run(x, frame_maps=[plan_spec_substitute(x.shape[0], max_t, num_t_sub)])
