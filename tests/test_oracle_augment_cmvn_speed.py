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
