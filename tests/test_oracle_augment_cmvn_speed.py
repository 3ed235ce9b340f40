"""Pins the augmentation / CMVN / speed / collate oracles to goldens made by the reference's own functions."""
import os
import random

import numpy as np
import pytest

from oracle import augment as A
from oracle import cmvn as C
from oracle import collate as K
from oracle import speed as S
from oracle import fbank as F


@pytest.fixture(scope='module')
def aug(golden_dir):
    return np.load(os.path.join(golden_dir, 'augment.npz'))


@pytest.mark.parametrize('i', range(5))
def test_spec_aug_sub_bit_exact(aug, i):
    x = aug['x%d' % i]
    random.seed(1000 + i)
    assert np.array_equal(A.spec_augmentation(x, 3, 2, 50, 10), aug['aug%d' % i])
    random.seed(2000 + i)
    assert np.array_equal(A.spec_substitute(x, max_t=30, num_t_sub=3), aug['sub%d' % i])
    random.seed(3000 + i)
    y = A.spec_augmentation(A.spec_substitute(x, max_t=30, num_t_sub=3), 3, 2, 50, 10)
    assert np.array_equal(y, aug['subaug%d' % i])


@pytest.mark.parametrize('i', range(5))
def test_substitute_is_an_index_map(aug, i):
    x = aug['x%d' % i]
    random.seed(2000 + i)
    subs = A.plan_spec_substitute(x.shape[0], max_t=30, num_t_sub=3)
    idx = A.substitute_index_map(x.shape[0], subs)
    assert np.array_equal(x[idx], aug['sub%d' % i])


@pytest.mark.parametrize('i', [0, 1, 2, 4])
def test_normalization(aug, i):
    assert np.array_equal(A.normalization(aug['x%d' % i]), aug['norm%d' % i])


def test_normalization_has_no_epsilon():
    with np.errstate(all='ignore'):
        y = A.normalization(np.full((5, 3), -15.9424, np.float32))
    assert np.isnan(y).all()


def test_speed_generator_quirk(aug):
    random.seed(5)
    draws = [S.speed_generator([0.9, 1.1, 0.1]) for _ in range(8)] + [S.speed_generator(None)] + [S.speed_generator([1.05])]
    assert np.array_equal(np.array(draws), aug['speed_draws'])
    assert all(d == 9 * 0.1 for d in draws[:9])
    random.seed(6)
    assert np.array_equal(np.array([S.speed_generator([0.9, 1.1, 0]) for _ in range(8)]), aug['speed_draws_uniform'])


def test_cmvn_loaders_and_apply(golden_dir):
    g = np.load(os.path.join(golden_dir, 'cmvn.npz'))
    mean, istd = C.load_cmvn(os.path.join(golden_dir, 'cmvn_stats.json'), True)
    assert np.array_equal(mean, g['mean_json']) and np.array_equal(istd, g['istd_json'])
    mean_k, istd_k = C.load_cmvn(os.path.join(golden_dir, 'cmvn_stats.kaldi.txt'), False)
    assert np.array_equal(mean_k, g['mean_kaldi']) and np.array_equal(istd_k, g['istd_kaldi'])
    assert np.array_equal(C.global_cmvn(g['x'], mean, istd), g['y'])
    assert np.array_equal(C.global_cmvn(g['x'], mean, istd, norm_var=False), g['y_novar'])
    # padded cells are (0 - mean) * istd after CMVN, not 0
    assert np.array_equal(g['y'][0, -1], ((np.float32(0) - mean.astype(np.float32)) * istd.astype(np.float32)))


def test_cmvn_stats_roundtrip(tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, 'cmvn.npz'))
    feats = [g['x'][0][:48], g['x'][1]]
    s, q, n = C.compute_cmvn_stats(feats)
    p = str(tmp_path / 'cmvn.json')
    C.write_json_cmvn(p, s, q, n)
    mean, istd = C.load_cmvn(p, True)
    allf = np.concatenate(feats).astype(np.float64)
    np.testing.assert_allclose(mean, allf.mean(0), rtol=1e-12)
    np.testing.assert_allclose(istd, 1.0 / allf.std(0), rtol=1e-9)


def test_speed_oracle_matches_torchaudio_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'speed.npz'))
    for s, tag in [(0.9, '090'), (1.1, '110')]:
        y = S.speed_perturb(g['x'], 16000, s)
        assert y.shape == g['y' + tag].shape
        assert y.shape[0] == S.output_length(8000, *S.speed_ratio(s))
        assert np.abs(y - g['y' + tag]).max() < 0.05            # int16 scale; stated resampler tolerance
        assert np.abs(F.fbank(y) - g['fb' + tag]).max() < 2e-3
    assert S.speed_ratio(0.9) == (9, 10) and S.speed_ratio(1.1) == (11, 10)
    x = g['x']
    assert S.speed_perturb(x, 16000, 1.0) is x                  # audio_processor.py:31 returns the input


def test_collate_port_matches_reference_golden(golden_dir, tables):
    g = np.load(os.path.join(golden_dir, 'collate.npz'))
    window, mel = tables
    lens = [16000, 9000, 5200, 12345, 300, 7777]
    batch = [('utt%d' % i, (g['pcm%d' % i].astype(np.float32), 16000), [i + 1] * (i + 2), 1.0)
             for i in range(len(lens))]
    batch.append(('seg', (g['pcm0'][4000:12000].astype(np.float32), 16000), [9, 9], 1.0))
    conf = {'resample_rate': 16000, 'speed_perturb_rate': 0, 'speeds': [0.9, 1.1, 0.1], 'wav_dither': 0.0,
            'mel_bins': 80}

    def fb(w, mel_bins, dither, sr):
        return F.fbank(w, num_mel_bins=mel_bins, window=window, mel=mel)

    for tag, kw in [('plain', dict(normalization=False)),
                    ('norm_aug', dict(normalization=True, spec_aug=True,
                                      spec_aug_conf=dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10))),
                    ('sub_aug', dict(normalization=False, spec_sub=True, spec_sub_conf=dict(num_t_sub=3, max_t=30),
                                     spec_aug=True, spec_aug_conf=dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10)))]:
        fn = K.AudioCollate(feature_extraction_conf=conf, fbank_fn=fb, **kw)
        random.seed(4242)
        keys, out = fn([batch])
        assert list(keys) == list(g[tag + '_keys'])              # sorted by length desc, utt4 dropped
        assert np.array_equal(out['features_length'], g[tag + '_features_length'])
        assert np.array_equal(out['targets'], g[tag + '_targets'])
        assert np.array_equal(out['targets_length'], g[tag + '_targets_length'])
        ref = g[tag + '_features']
        assert out['features'].shape == ref.shape and out['features'].dtype == np.float32
        assert np.array_equal(out['features'] == 0, ref == 0)    # masks and padding bit-exact
        assert np.abs(out['features'] - ref).max() < 2e-3


def test_soxlike_stand_in_meets_its_specification():
    """The second resampler oracle (oracle/speed.py: soxlike_kernel) is a design, not a port: check the design.  Pass band
    (<= 0.95 of the lower Nyquist): a tone comes out as the analytically resampled tone to 1e-5; stop band (>= the lower
    Nyquist): >= 120 dB down; DC gain 1."""
    for sp in (0.9, 1.1):
        orig, new = S.speed_ratio(sp)
        k, width = S.soxlike_kernel(orig, new, dtype=np.float64)
        assert k.shape == (new, 2 * width + orig)
        n = 6000
        t = np.arange(n, dtype=np.float64)
        fn = 0.5 * min(1.0, new / orig)
        for f_rel, passband in ((0.05, True), (0.5, True), (0.94, True), (1.02, False), (1.5, False)):
            f = f_rel * fn                                       # cycles per input sample
            if f >= 0.5:
                continue
            x = np.cos(2 * np.pi * f * t + 0.3)
            y = S.resample(x, orig, new, dtype=np.float64, kernel=k)
            m = np.arange(len(y), dtype=np.float64) * orig / new  # output instants in input samples
            mid = slice(600, len(y) - 600)
            if passband:
                assert np.abs(y[mid] - np.cos(2 * np.pi * f * m[mid] + 0.3)).max() < 1e-5
            else:
                assert np.abs(y[mid]).max() < 10 ** (-120 / 20.0)
        assert abs(k.sum() / new - 1.0) < 1e-6


def test_measured_distance_torchaudio_sinc_vs_soxlike():
    """How far the product's resampler (torchaudio's width-6 hann sinc, the substitute the GPU path is pinned to) is
    from a sox-quality one, on the log-mel features -- the numbers quoted in DESIGN.md section 2.  Below ~6.5 kHz the
    two agree closely; the top mel bins (inside / above the transition band of the shorter filter) differ by design."""
    from oracle import signals
    worst_low, worst_top = 0.0, 0.0
    for kind in ('speech', 'white'):
        w = signals.make(kind, 48000, 11).astype(np.float64)
        for sp in (0.9, 1.1):
            a = S.speed_perturb(w, 16000, sp, dtype=np.float64)
            b = S.speed_perturb_soxlike(w, 16000, sp)
            assert len(a) == len(b)
            d = np.abs(F.fbank(a, dtype=np.float64) - F.fbank(b, dtype=np.float64))
            worst_low = max(worst_low, float(d[:, :60].max()))
            worst_top = max(worst_top, float(d[:, 60:].max()))
    assert worst_low < 0.6           # measured 0.52 (speech-like 1/f spectrum, speed 0.9: bins ~60 dB below the frame's peak); white noise: 0.07
    assert 1.0 < worst_top < 25.0    # measured 19.2: speed 0.9 leaves 7.2-8 kHz empty with sox, images with the short sinc


def test_oracle_resample_pinned_on_long_ratios(golden_dir):
    """The oracle's polyphase restatement (kernel evaluated in float64) against torchaudio.functional.resample goldens
    for 441:160 and 953:1000.  torchaudio builds the kernel on an fp32 grid (p/new + idx/orig: a difference of two
    numbers near 1), which is noisy for long periods: measured 0.07 and 0.79 on the int16 scale (amplitude 1.7e4, i.e.
    5e-5 relative) -- the bound below states that gap, it is torchaudio's rounding, not a resampler difference."""
    g = np.load(os.path.join(golden_dir, 'resample_long.npz'))
    for name, (o, n), tol in (('441_160', (441, 160), 0.1), ('953_1000', (953, 1000), 1.0)):
        y = S.resample(g['x_' + name], o, n, dtype=np.float64)
        assert y.shape == g['y_' + name].shape and np.abs(y - g['y_' + name]).max() <= tol
