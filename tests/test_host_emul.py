"""The in-register FFT templates of the CUDA kernel (csrc/oe_fft.h), compiled for the host and driven through
the same 16-thread / two-frame (packed f32x2, emulated as a float pair) decomposition, against numpy (CPU)."""
import ctypes

import numpy as np
import pytest

from openeat_b200 import _lib
from oracle import fbank as F
from oracle import signals


@pytest.fixture(scope='module', params=['gen1', 'gen2'])
def emul(request):
    """gen1: oe_fbank_kernel's decomposition (radix-2 stages, full row exchange, one untangle per bin);
    gen2: oe_fbank2_kernel's (radix-4 stages, half-row partner exchange, pair untangle, lane-0 / bin-128 rules).
    gen2 leaves the weightless bins 0 and 256 undefined; they are zeroed here before the comparison."""
    _lib.build()
    lib = ctypes.CDLL(_lib.EMUL_PATH)
    P = ctypes.POINTER(ctypes.c_float)
    fn = lib.oe_emul_frame_pair if request.param == 'gen1' else lib.oe_emul_frame_pair_v2
    fn.argtypes = [P, P, P, P]

    def power(ha, hb):
        ha = np.ascontiguousarray(ha, dtype=np.float32)
        hb = np.ascontiguousarray(hb, dtype=np.float32)
        pa, pb = np.zeros(257, np.float32), np.zeros(257, np.float32)
        fn(ha.ctypes.data_as(P), hb.ctypes.data_as(P), pa.ctypes.data_as(P), pb.ctypes.data_as(P))
        return pa[:256], pb[:256]
    power.gen2 = request.param == 'gen2'
    return power


def ref_power(h):
    return (np.abs(np.fft.rfft(h.astype(np.float64), 512)) ** 2)[:256]


@pytest.mark.parametrize('kind', signals.CLASSES)
def test_frame_power_spectrum(emul, kind, tables):
    x = signals.make(kind, 400 + 160 * 5, 3).astype(np.float32)
    h = F.windowed_frames(x, np.float32, window=tables[0])[:, :400]
    for a, b in zip(h[0::2], h[1::2]):
        for got, row in zip(emul(a, b), (a, b)):
            ref = ref_power(row)
            assert np.abs(got[1:] - ref[1:]).max() <= 2e-6 * max(ref.max(), 1e-30)


def test_the_two_packed_frames_do_not_interact(emul):
    """A loud and a -100 dB frame side by side: each must be as accurate as on its own (the two frames are the
    two halves of an f32x2 register, never mixed -- unlike packing them as real/imaginary parts)."""
    rng = np.random.default_rng(5)
    loud = (rng.normal(0, 20000, 400) * np.hanning(400)).astype(np.float32)
    quiet = (rng.normal(0, 0.2, 400) * np.hanning(400)).astype(np.float32)
    pa, pb = emul(loud, quiet)
    qa, qb = emul(quiet, loud)
    assert np.array_equal(pa[1:], qb[1:]) and np.array_equal(pb[1:], qa[1:])
    ref = ref_power(quiet)
    assert np.abs(pb[1:] - ref[1:]).max() <= 2e-6 * ref.max()


def test_impulse_and_tone_bins(emul):
    h = np.zeros(400, np.float32)
    h[3] = 1.0
    pa, pb = emul(h, 2 * h)
    assert np.allclose(pa[1:], 1.0, atol=1e-6) and np.allclose(pb[1:], 4.0, atol=4e-6)      # flat spectra
    n = np.arange(400)
    for k in (1, 8, 15, 16, 37, 128, 200, 255):                    # every row class of the 16 x 16 split
        pa, pb = emul(np.cos(2 * np.pi * k * n / 512).astype(np.float32), h)
        assert pa[1:].argmax() + 1 == k
