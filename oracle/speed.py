"""Substitute speed-perturb oracle (tests only).

The reference's ``_speed_perturb`` (``openeat/dataset/audio_processor.py:19-35``)
runs libsox ``speed <s>`` + ``rate <sr>`` through ``torchaudio.sox_effects`` --
libsox is not installed and that module no longer exists in torchaudio 2.x, so the
real chain cannot run here: PARITY AGAINST LIBSOX IS UNPINNED.

The substitute restated below is ``torchaudio.functional.speed`` (torchaudio
2.11.0 ``functional/functional.py:2385-2423`` -> ``resample`` ``:1435-1490`` ->
``_get_sinc_resample_kernel`` ``:1305-1398`` / ``_apply_sinc_resample_kernel``
``:1401-1432``): hann-windowed sinc, ``lowpass_filter_width=6``, ``rolloff=0.99``.
It has the same semantics (relabel the rate to ``sr*speed``, resample back to
``sr``; output length ~ N/speed) and is pinned against live torchaudio in
``tests/test_oracle_speed.py``.

``_speed_generator`` (``audio_processor.py:5-18``) is restated verbatim in
behaviour, including its quirk of always returning ``speeds[0]`` when the step is
non-zero.
"""
import math
import random

import numpy as np


def speed_generator(speeds, rng=random):
    """audio_processor.py:5-18."""
    if speeds is None:
        speeds = [0.9, 1.1, 0.1]
    speeds = [float(s) for s in speeds]
    if len(speeds) > 1:
        assert speeds[1] > speeds[0], 'speeds is wrong !'
        if speeds[2] != 0:
            speed = rng.randrange(int(speeds[0] / speeds[2]), int(speeds[0] / speeds[2]) + 1)
            speed *= speeds[2]
        else:
            speed = speeds[0] + rng.random() * (speeds[1] - speeds[0])
    else:
        speed = speeds[0]
    return speed


def speed_ratio(speed, sample_rate=16000):
    """functional.py:2408-2413: (orig, new) after gcd reduction.  0.9 -> (9, 10), 1.1 -> (11, 10)."""
    src = int(speed * sample_rate)
    dst = int(sample_rate)
    g = math.gcd(src, dst)
    return src // g, dst // g


def sinc_resample_kernel(orig, new, lowpass_filter_width=6, rolloff=0.99, dtype=np.float32):
    """functional.py:1343-1398 (sinc_interp_hann).  Returns (kernel[new, 2*width+orig], width).
    Evaluated in float64 and rounded once (torch evaluates the grid in the waveform dtype)."""
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t = t * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / orig
    with np.errstate(invalid='ignore', divide='ignore'):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * scale
    return k.astype(dtype), width


def output_length(n, orig, new):
    """functional.py:1427: ceil(new * n / orig)."""
    return int(math.ceil(new * n / orig))


def resample(wave, orig, new, dtype=np.float32, kernel=None):
    """functional.py:1401-1432: pad (width, width+orig), strided correlation, interleave
    the ``new`` phases, truncate to ceil(new*N/orig)."""
    x = np.asarray(wave, dtype=dtype).reshape(-1)
    if orig == new:
        return x
    if kernel is None:
        k, width = sinc_resample_kernel(orig, new, dtype=dtype)
    else:
        k = np.asarray(kernel, dtype=dtype)
        width = (k.shape[1] - orig) // 2
    n = x.shape[0]
    xp = np.concatenate([np.zeros(width, dtype), x, np.zeros(width + orig, dtype)])
    taps = k.shape[1]
    m = (xp.shape[0] - taps) // orig + 1
    idx = np.arange(m)[:, None] * orig + np.arange(taps)[None, :]
    seg = xp[idx]                                   # (m, taps)
    y = seg @ k.T                                   # (m, new)
    return y.reshape(-1)[:output_length(n, orig, new)].astype(dtype)


def speed_perturb(wave, sample_rate, speed, dtype=np.float32):
    """Substitute for audio_processor.py:19-35.  Returns the input unchanged for speed == 1.0."""
    if speed == 1.0:
        return wave
    orig, new = speed_ratio(speed, sample_rate)
    return resample(wave, orig, new, dtype=dtype)


# ---------------------------------------------------------------------------------------------------
# Second, independent resampler oracle: a stand-in for libsox `rate` at its default quality.
#
# sox(1), `rate`: the default quality is -h ("high"): 95 % of the band preserved, 125 dB rejection, linear phase, and
# without -a the stop band begins at the Nyquist frequency (no aliasing / imaging).  libsox is not available here, so
# its multi-stage implementation cannot be run or pinned; what CAN be stated independently of any implementation is
# the specification, and any linear-phase filter meeting it agrees with libsox's to ~1e-6 in the pass band -- the two
# can only differ inside the 5 % transition band.  `soxlike_kernel` designs such a filter directly (Kaiser-windowed
# sinc, pass-band edge 0.95 fn, stop-band edge fn, fn = the lower of the two Nyquist frequencies) in the polyphase
# layout `resample` takes.  It is used to MEASURE how far the product's resampler (torchaudio's width-6 hann sinc, the
# substitute oracle above) is from a sox-quality one (tests/test_oracle_augment_cmvn_speed.py);
# parity against libsox itself stays unpinned.
def soxlike_kernel(orig, new, passband=0.95, rejection_db=125.0, dtype=np.float32):
    """Returns (kernel[new, 2*width+orig], width): y[m*new + p] = sum_q kernel[p][q] * xpad[m*orig + q]."""
    fn = 0.5 * min(1.0, new / orig)                 # cycles per INPUT sample
    f_pass, f_stop = passband * fn, fn
    delta = f_stop - f_pass
    beta = 0.1102 * (rejection_db - 8.7)            # Kaiser's formulas
    half = int(math.ceil((rejection_db - 7.95) / (14.36 * delta) / 2.0))    # half length in input samples
    width = half
    fc = 0.5 * (f_pass + f_stop)
    q = np.arange(-width, width + orig, dtype=np.float64)[None, :]
    p = np.arange(new, dtype=np.float64)[:, None]
    t = q - p * orig / new                          # input-sample distance from the output instant
    with np.errstate(invalid='ignore', divide='ignore'):
        h = np.where(t == 0, 2.0 * fc, np.sin(2.0 * math.pi * fc * t) / (math.pi * t))
    r = np.clip(1.0 - (t / half) ** 2, 0.0, None)
    w = np.where(np.abs(t) <= half, np.i0(beta * np.sqrt(r)) / np.i0(beta), 0.0)
    return (h * w).astype(dtype), width


def speed_perturb_soxlike(wave, sample_rate, speed, dtype=np.float64):
    """`speed <s>` + `rate <sr>` with the sox-quality stand-in filter (see above)."""
    if speed == 1.0:
        return np.asarray(wave, dtype=dtype)
    orig, new = speed_ratio(speed, sample_rate)
    k, _ = soxlike_kernel(orig, new, dtype=dtype)
    return resample(wave, orig, new, dtype=dtype, kernel=k)
