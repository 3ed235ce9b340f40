"""Deterministic synthetic test signals (tests / benchmarks only).

All signals are int16 (the on-disk PCM dtype; the reference turns them into the
same integer values as fp32 at ``openeat/dataset/dataset.py:75``).  Classes follow
SURVEY.md section 8c: white noise alone hides dynamic-range errors.
"""
import hashlib

import numpy as np

CLASSES = ('white', 'speech', 'lsb', 'zero', 'dcsine', 'square')


def make(kind, n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    if kind == 'white':                       # Gaussian sigma=3000, clipped
        x = rng.normal(0.0, 3000.0, n)
    elif kind == 'speech':                    # 1/f-shaped noise with a 4 Hz amplitude envelope
        spec = np.fft.rfft(rng.normal(0.0, 1.0, n))
        f = np.arange(spec.shape[0], dtype=np.float64)
        f[0] = 1.0
        x = np.fft.irfft(spec / f, n)
        x = x / (np.abs(x).max() + 1e-12)
        env = 0.55 + 0.45 * np.sin(2 * np.pi * 4.0 * t / 16000.0 + rng.uniform(0, 6.28))
        x = 20000.0 * x * env
    elif kind == 'lsb':                       # +-1 LSB noise
        x = rng.integers(-1, 2, n).astype(np.float64)
    elif kind == 'zero':
        x = np.zeros(n)
    elif kind == 'dcsine':                    # 440 Hz sine on a DC offset
        x = 10000.0 * np.sin(2 * np.pi * 440.0 * t / 16000.0) + 3000.0
    elif kind == 'square':                    # full-scale 100 Hz square
        x = np.where(np.sin(2 * np.pi * 100.0 * t / 16000.0) >= 0, 32767.0, -32768.0)
    else:
        raise ValueError(kind)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def digest(x):
    return hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()


def lengths_uniform(count, lo_s, hi_s, seed, sample_rate=16000):
    """Utterance lengths in samples, uniform in [lo_s, hi_s] seconds."""
    rng = np.random.default_rng(seed)
    return np.round(rng.uniform(lo_s, hi_s, count) * sample_rate).astype(np.int64)
