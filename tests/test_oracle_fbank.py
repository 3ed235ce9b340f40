"""Pins the numpy fbank oracle to the torchaudio-generated golden vectors (CPU)."""
import os

import numpy as np
import pytest

from oracle import fbank as F
from oracle import signals

CASES = ['fbank_white_400', 'fbank_white_559', 'fbank_white_560', 'fbank_white_8000', 'fbank_white_80000',
         'fbank_white_560000', 'fbank_speech_8000', 'fbank_speech_80000', 'fbank_lsb_8000', 'fbank_zero_8000',
         'fbank_dcsine_8000', 'fbank_square_8000']


def test_tables_close_to_torch(tables):
    window, mel = tables
    assert window[0] == 0.0 and window[-1] == 0.0
    assert np.abs(F.povey_window() - window).max() < 5e-7
    own = F.mel_banks()
    assert own.shape == mel.shape == (80, 256)
    assert np.abs(own - mel).max() < 5e-5
    assert (mel != 0).sum() == 501 and (mel[:, 0] == 0).all()
    assert ((mel != 0).sum(axis=0) <= 2).all()


def test_frame_count_formula():
    assert [F.num_frames(n) for n in (399, 400, 559, 560, 8000, 80000, 560000)] == [0, 1, 1, 2, 48, 498, 3498]


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_golden(name, golden_dir, manifest, tables):
    meta = manifest[name]
    x = signals.make(meta['kind'], meta['samples'], meta['seed'])
    assert signals.digest(x) == meta['sha256'], 'synthetic signal generator drifted'
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    stride = int(g['stride'])
    window, mel = tables
    o32 = F.fbank(x.astype(np.float32), window=window, mel=mel)
    o64 = F.fbank(x.astype(np.float64), dtype=np.float64)
    assert o32.shape == (meta['frames'], 80)
    # fp64 restatement vs torchaudio fp64: table ulps only
    assert np.abs(o64[::stride] - g['y64']).max() < 2e-4
    np.testing.assert_allclose(o64.sum(axis=0), g['colsum64'], rtol=0, atol=2e-4 * meta['frames'])
    # fp32 restatement vs torchaudio fp32: two fp32 FFTs differ by about the fp32-vs-fp64 gap
    tol = max(1e-3, 2.0 * meta['gap32_64'])
    assert np.abs(o32[::stride] - g['y32']).max() < tol
    assert np.abs(o32[::stride] - g['y64']).max() < tol


def test_zero_signal_is_log_eps(golden_dir):
    g = np.load(os.path.join(golden_dir, 'fbank_zero_8000.npz'))
    assert np.all(g['y32'] == np.log(F.EPS_F32))
    assert np.all(F.fbank(np.zeros(8000, np.float32)) == np.log(F.EPS_F32))


def test_short_utterance_raises(manifest):
    assert manifest['short_399_raises'] is True
    with pytest.raises(AssertionError):
        F.fbank(np.zeros(399, np.float32))


def test_frames_are_independent():
    """fbank over 16-frame windows == fbank over the stream."""
    x = signals.make('speech', 160 * 63 + 400, 3).astype(np.float32)
    whole = F.fbank(x)
    parts = [F.fbank(x[160 * 16 * i: 160 * 16 * i + 160 * 15 + 400]) for i in range(4)]
    assert np.array_equal(np.concatenate(parts), whole)
