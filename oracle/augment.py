"""Numpy restatement of ``openeat/dataset/feature_processor.py`` (oracle; tests only).

The reference draws its random indices from Python's global ``random`` module while
it mutates the array.  Here each processor is split into ``plan_*`` (makes exactly
the same ``random`` calls, in the same order, and returns the indices) and
``apply_*`` (pure index work), so the CUDA path can be checked bit-exactly against
the very same plan.  ``plan`` followed by ``apply`` is the reference function.
"""
import random

import numpy as np

SPEC_MASK = 0  # feature_processor.py:4


def normalization(feature):
    """feature_processor.py:5-8: per-utterance, per-bin (x - mean) / std, ddof 0, no eps."""
    mean = np.mean(feature, axis=0)
    std = np.std(feature, axis=0)
    return (feature - mean) / std


def plan_spec_augmentation(num_frames, num_freq, num_t_mask=2, num_f_mask=2, max_t=50, max_f=10,
                           rng=random):
    """RNG calls of feature_processor.py:31-41.  Returns (t_masks, f_masks), lists of
    half-open [start, end) ranges already clipped to the array."""
    t_masks, f_masks = [], []
    for _ in range(num_t_mask):
        start = rng.randint(0, num_frames - 1)
        length = rng.randint(1, max_t)
        t_masks.append((start, min(num_frames, start + length)))
    for _ in range(num_f_mask):
        start = rng.randint(0, num_freq - 1)
        length = rng.randint(1, max_f)
        f_masks.append((start, min(num_freq, start + length)))
    return t_masks, f_masks


def apply_spec_augmentation(x, t_masks, f_masks):
    """feature_processor.py:27-42 with the indices given."""
    y = np.copy(x)
    for s, e in t_masks:
        y[s:e, :] = SPEC_MASK
    for s, e in f_masks:
        y[:, s:e] = SPEC_MASK
    return y


def spec_augmentation(x, num_t_mask=2, num_f_mask=2, max_t=50, max_f=10, rng=random):
    """feature_processor.py:10-42."""
    t, f = plan_spec_augmentation(x.shape[0], x.shape[1], num_t_mask, num_f_mask, max_t, max_f, rng)
    return apply_spec_augmentation(x, t, f)


def plan_spec_substitute(num_frames, max_t=20, num_t_sub=3, rng=random):
    """RNG calls of feature_processor.py:57-63.  Returns [(start, end, pos), ...]."""
    subs = []
    for _ in range(num_t_sub):
        start = rng.randint(0, num_frames - 1)
        length = rng.randint(1, max_t)
        end = min(num_frames, start + length)
        pos = rng.randint(0, start)
        subs.append((start, end, pos))
    return subs


def apply_spec_substitute(x, subs):
    """feature_processor.py:55-64 with the indices given (numpy slice assignment is
    overlap-safe: the right-hand side is read before anything is written)."""
    y = np.copy(x)
    for start, end, pos in subs:
        y[start:end, :] = y[start - pos:end - pos, :].copy()
    return y


def spec_substitute(x, max_t=20, num_t_sub=3, rng=random):
    """feature_processor.py:44-64."""
    return apply_spec_substitute(x, plan_spec_substitute(x.shape[0], max_t, num_t_sub, rng))


def substitute_index_map(num_frames, subs):
    """The composed frame-index map of a substitution plan: ``y[t] == x[idx[t]]``.
    Because every substitution copies whole rows, the sequence of copies is a
    composition of index maps."""
    idx = np.arange(num_frames, dtype=np.int32)
    for start, end, pos in subs:
        idx[start:end] = idx[start - pos:end - pos].copy()
    return idx
