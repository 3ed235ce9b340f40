"""Numpy restatement of ``torchaudio.compliance.kaldi.fbank`` (oracle; tests only).

Follows torchaudio 2.11.0 ``compliance/kaldi.py`` with the keyword arguments the
reference passes at ``openeat/dataset/dataset.py:93-100``:
``num_mel_bins=mel_bins, frame_length=25, frame_shift=10, dither=wav_dither,
energy_floor=0.0, sample_frequency=sr`` and every other argument at its default
(povey window, preemphasis 0.97, remove_dc_offset, snip_edges, round to power of
two, low_freq 20, high_freq 0 (= Nyquist), use_power, use_log_fbank, no energy,
no vtln).

``dtype=np.float32`` mirrors the fp32 path the reference runs; ``np.float64`` is
the high-precision oracle used to judge tonal signals.
"""
import math

import numpy as np

EPS_F32 = np.float32(1.1920928955078125e-07)  # torch.finfo(torch.float).eps, kaldi.py:31-37


def next_power_of_2(x):
    """kaldi.py:39-41."""
    return 1 if x == 0 else 2 ** (x - 1).bit_length()


def window_properties(sample_frequency=16000.0, frame_length=25.0, frame_shift=10.0):
    """kaldi.py:138-140: (shift, size, padded) in samples."""
    shift = int(sample_frequency * frame_shift * 0.001)
    size = int(sample_frequency * frame_length * 0.001)
    return shift, size, next_power_of_2(size)


def num_frames(num_samples, window_size=400, window_shift=160):
    """kaldi.py:63-67 (snip_edges=True). 0 means the reference drops the utterance
    (assert at kaldi.py:142 raises, caught at dataset.py:108-111)."""
    if num_samples < window_size:
        return 0
    return 1 + (num_samples - window_size) // window_shift


def povey_window(window_size=400, dtype=np.float32):
    """kaldi.py:98-100: hann(window_size, periodic=False) ** 0.85.

    Evaluated in float64 and rounded once to ``dtype``; the test-suite checks it
    against torch's fp32 table (<= 2 ulp).  w[0] == w[-1] == 0 exactly.
    """
    j = np.arange(window_size, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * math.pi * j / (window_size - 1))
    hann[0] = 0.0
    hann[-1] = 0.0
    return (hann ** 0.85).astype(dtype)


def mel_scale(freq):
    """kaldi.py:321-326."""
    return 1127.0 * np.log(1.0 + freq / 700.0)


def mel_banks(num_bins=80, padded=512, sample_freq=16000.0, low_freq=20.0, high_freq=0.0,
              dtype=np.float32):
    """kaldi.py:436-511 with vtln_warp_factor == 1.0.  Returns (num_bins, padded//2).

    torch evaluates the scalar mel limits in Python float64 and the per-bin /
    per-fft-bin tensors in fp32 (then ``.to(dtype)``); the same mixed precision is
    kept here.  numpy's ``logf`` and torch's differ by an ulp on some fft bins,
    which the division by the ~35-mel bin width turns into <= 2e-5 absolute weight
    differences -- tests therefore also run the oracle with torch's exact tables
    (stored in tests/golden/tables.npz) passed through ``window=`` / ``mel=``.
    """
    num_fft_bins = padded // 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)

    f32 = np.float32  # torch builds this table in fp32 whatever the waveform dtype (kaldi.py:621-624)
    b = np.arange(num_bins, dtype=np.int64)[:, None]
    # torch: python-float + int64 tensor * python-float -> float32 tensors
    left = (f32(mel_low) + b.astype(f32) * f32(delta)).astype(f32)
    center = (f32(mel_low) + (b.astype(f32) + f32(1.0)) * f32(delta)).astype(f32)
    right = (f32(mel_low) + (b.astype(f32) + f32(2.0)) * f32(delta)).astype(f32)
    k = np.arange(num_fft_bins, dtype=f32)
    mel = (f32(1127.0) * np.log(f32(1.0) + (f32(fft_bin_width) * k) / f32(700.0))).astype(f32)[None, :]
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return np.maximum(f32(0.0), np.minimum(up, down)).astype(dtype)


def frames_of(wave, window_size=400, window_shift=160):
    """kaldi.py:44-83 (snip_edges=True): (m, window_size) strided view, copied."""
    m = num_frames(wave.shape[0], window_size, window_shift)
    if m == 0:
        return np.zeros((0, window_size), dtype=wave.dtype)
    idx = np.arange(m)[:, None] * window_shift + np.arange(window_size)[None, :]
    return wave[idx]


def windowed_frames(wave, dtype=np.float32, window_size=400, window_shift=160,
                    preemph=0.97, window=None):
    """kaldi.py:183-211: DC removal, pre-emphasis (replicate pad), window, zero pad."""
    x = np.asarray(wave, dtype=dtype)
    f = frames_of(x, window_size, window_shift)
    if f.shape[0] == 0:
        return np.zeros((0, next_power_of_2(window_size)), dtype=dtype)
    mean = (np.sum(f, axis=1, dtype=dtype) / dtype(window_size))[:, None]
    f = f - mean
    prev = np.concatenate([f[:, :1], f[:, :-1]], axis=1)
    f = f - dtype(preemph) * prev
    w = povey_window(window_size, dtype) if window is None else np.asarray(window, dtype=dtype)
    f = f * w[None, :]
    padded = next_power_of_2(window_size)
    out = np.zeros((f.shape[0], padded), dtype=dtype)
    out[:, :window_size] = f
    return out


def fbank(wave, num_mel_bins=80, sample_frequency=16000.0, frame_length=25.0, frame_shift=10.0,
          dtype=np.float32, window=None, mel=None):
    """Log-mel filterbank of a mono waveform (values on the int16 scale, as the
    reference multiplies by 1<<15 at dataset.py:75).  Returns (m, num_mel_bins).

    Raises AssertionError when the waveform is shorter than one window, like
    kaldi.py:142 -- the caller (dataset.py:108-111) drops such utterances.
    """
    shift, size, padded = window_properties(sample_frequency, frame_length, frame_shift)
    wave = np.asarray(wave).reshape(-1)
    assert 2 <= size <= wave.shape[0], "choose a window size %d that is [2, %d]" % (size, wave.shape[0])
    h = windowed_frames(wave, dtype, size, shift, 0.97, window)
    spec = np.fft.rfft(h, axis=1)                       # kaldi.py:616
    if dtype == np.float32:
        spec = spec.astype(np.complex64)
    power = np.abs(spec).astype(dtype) ** dtype(2.0)    # kaldi.py:616-618 (abs then pow)
    if mel is None:
        mel = mel_banks(num_mel_bins, padded, sample_frequency, 20.0, 0.0, dtype)
    mel = np.asarray(mel, dtype=dtype)
    melp = np.concatenate([mel, np.zeros((mel.shape[0], 1), dtype=dtype)], axis=1)  # kaldi.py:627
    energies = power @ melp.T                           # kaldi.py:630
    eps = dtype(EPS_F32)  # kaldi.py:31-37: EPSILON is finfo(float32).eps cast to the waveform dtype
    return np.log(np.maximum(energies, eps)).astype(dtype)  # kaldi.py:631-633
