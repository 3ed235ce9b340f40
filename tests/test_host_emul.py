"""The in-register FFT templates of the CUDA kernel (csrc/oe_fft.h), compiled for the host and
driven through the same 16-thread decomposition, against numpy (CPU, no GPU needed)."""
import ctypes

import numpy as np
import pytest

from openeat_b200 import _lib
from oracle import fbank as F
from oracle import signals


@pytest.fixture(scope='module')
def emul():
    _lib.build()
    lib = ctypes.CDLL(_lib.EMUL_PATH)
    P = ctypes.POINTER(ctypes.c_float)
    lib.oe_emul_frame.argtypes = [P, P]

    def power(h):
        h = np.ascontiguousarray(h, dtype=np.float32)
        pw = np.zeros(257, np.float32)
        lib.oe_emul_frame(h.ctypes.data_as(P), pw.ctypes.data_as(P))
        return pw
    return power


@pytest.mark.parametrize('kind', signals.CLASSES)
def test_frame_power_spectrum(emul, kind, tables):
    x = signals.make(kind, 400 + 160 * 5, 3).astype(np.float32)
    h = F.windowed_frames(x, np.float32, window=tables[0])[:, :400]
    for row in h:
        ref = np.abs(np.fft.rfft(row.astype(np.float64), 512)) ** 2
        got = emul(row)
        scale = max(ref.max(), 1e-30)
        assert np.abs(got - ref).max() <= 2e-6 * scale


def test_impulse_and_tone_bins(emul):
    h = np.zeros(400, np.float32)
    h[3] = 1.0
    assert np.allclose(emul(h), 1.0, atol=1e-6)                 # flat spectrum
    n = np.arange(400)
    for k in (1, 8, 16, 37, 128, 200, 255):                      # every row class of the 16x16 split
        pw = emul(np.cos(2 * np.pi * k * n / 512).astype(np.float32))
        assert pw.argmax() == k
