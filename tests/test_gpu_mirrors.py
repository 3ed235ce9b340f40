"""GPU tests of the reference-facing Python API (same names / arguments / error behaviour as OpenEAT's)
against goldens produced by the reference's own functions (tests/golden, oracle/make_golden.py)."""
import os
import random
import wave

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import collate as K          # noqa: E402
from oracle import fbank as F            # noqa: E402
from oracle import signals               # noqa: E402

CONF = {'resample_rate': 16000, 'speed_perturb_rate': 0, 'speeds': [0.9, 1.1, 0.1], 'wav_dither': 0.0,
        'mel_bins': 80}
LENS = [16000, 9000, 5200, 12345, 300, 7777]
VARIANTS = [('plain', dict(normalization=False)),
            ('norm_aug', dict(normalization=True, spec_aug=True,
                              spec_aug_conf=dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10))),
            ('sub_aug', dict(normalization=False, spec_sub=True, spec_sub_conf=dict(num_t_sub=3, max_t=30),
                             spec_aug=True, spec_aug_conf=dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10)))]


def write_wav(path, pcm, sr=16000):
    with wave.open(str(path), 'wb') as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.asarray(pcm, dtype='<i2').tobytes())


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'collate.npz'))


@pytest.fixture()
def wav_batch(gold, tmp_path):
    """The very batch the reference's audio_collate_func saw: wav files on disk, one < 400 samples,
    plus a segmented entry ``path,start,end``."""
    batch = []
    for i in range(len(LENS)):
        p = tmp_path / ('u%d.wav' % i)
        write_wav(p, gold['pcm%d' % i])
        batch.append(('utt%d' % i, str(p), [i + 1] * (i + 2), 1.0))
    batch.append(('seg', str(tmp_path / 'u0.wav') + ',0.25,0.75', [9, 9], 1.0))
    return batch


@pytest.mark.parametrize('tag,kw', VARIANTS)
def test_audio_collate_func_matches_reference_output(gold, wav_batch, tag, kw):
    from openeat_b200.dataset import audio_collate_func
    fn = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, **kw)
    random.seed(4242)
    keys, out = fn([wav_batch])                                   # DataLoader wraps the pre-built batch in a list
    assert list(keys) == list(gold[tag + '_keys'])                # length-descending order, utt4 dropped
    feats = out['features'].cpu().numpy()
    ref = gold[tag + '_features']
    assert out['features'].dtype == torch.float32 and out['features_length'].dtype == torch.int32
    assert out['targets'].dtype == torch.int32 and out['targets_length'].dtype == torch.int32
    assert np.array_equal(out['features_length'].cpu().numpy(), gold[tag + '_features_length'])
    assert np.array_equal(out['targets'].cpu().numpy(), gold[tag + '_targets'])          # padded with -1
    assert np.array_equal(out['targets_length'].cpu().numpy(), gold[tag + '_targets_length'])
    assert feats.shape == ref.shape
    assert np.array_equal(feats == 0, ref == 0)                   # SpecAug / SpecSub indices and padding bit-exact
    assert np.abs(feats - ref).max() < 2e-3
    # CPU tensors on request, like the reference's DataLoader workers return
    fn_cpu = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, output_device='cpu', **kw)
    random.seed(4242)
    _, out_cpu = fn_cpu(wav_batch)                                # un-wrapped list form (dataset.py:186-187)
    assert not out_cpu['features'].is_cuda and np.array_equal(out_cpu['features'].numpy(), feats)


def test_extract_feature_signature_and_drop_convention(gold, wav_batch, capsys):
    from openeat_b200.dataset import _extract_feature
    keys, feats, labels = _extract_feature(wav_batch, CONF)
    assert 'choose a window size 400 that is [2, 300]' in capsys.readouterr().out   # printed, not raised
    assert keys == list(gold['plain_keys'])
    assert [f.shape[0] for f in feats] == gold['plain_features_length'].tolist()
    assert all(isinstance(f, np.ndarray) and f.dtype == np.float32 and f.shape[1] == 80 for f in feats)
    for f, n in zip(feats, gold['plain_features_length']):
        idx = keys.index(keys[[x.shape[0] for x in feats].index(f.shape[0])])
        assert np.abs(f - gold['plain_features'][idx, :n]).max() < 1e-3
    assert [len(l) for l in labels] == gold['plain_targets_length'].tolist()
    assert _extract_feature([('bad', '/nonexistent.wav', [1], 1.0)], CONF) == ([], [], [])   # never raises
    assert _extract_feature([], CONF) == ([], [], [])


def test_empty_batch_like_reference(gold):
    from openeat_b200.dataset import audio_collate_func
    fn = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, normalization=False)
    # a list of length 1 is unwrapped (dataset.py:186-187: the DataLoader hands over [pre-built batch])
    keys, out = fn([[('short', gold['pcm4'], [1, 2], 1.0)]])      # only a 300-sample utterance: dropped
    assert keys == [] and out['features'].numel() == 0 and out['features_length'].numel() == 0
    keys, out = fn([('short', gold['pcm4'], [1, 2], 1.0), ('bad', '/nonexistent.wav', [3], 1.0)])
    assert keys == [] and out['targets'].numel() == 0


def test_online_speed_perturb_in_collate(gold):
    """speed_perturb_rate=1 -> the reference's generator always yields 0.9 (appendix A.1); offline speeds
    arrive as item[3].  Checked against the CPU port with the substitute resampler oracle."""
    from openeat_b200.dataset import audio_collate_func
    conf = dict(CONF, speed_perturb_rate=1.0)
    batch = [('u%d' % i, gold['pcm%d' % i], [1, 2, 3], 1.0) for i in (0, 1, 2, 3)]
    random.seed(7)
    keys, out = audio_collate_func(data_type='wav', feature_extraction_conf=conf, normalization=False)(batch)
    obatch = [(k, (gold['pcm%d' % i].astype(np.float32), 16000), l, s) for (k, _, l, s), i in zip(batch, (0, 1, 2, 3))]
    random.seed(7)
    okeys, oout = K.AudioCollate(feature_extraction_conf=conf, normalization=False)(obatch)
    assert keys == okeys
    assert np.array_equal(out['features_length'].cpu().numpy(), oout['features_length'])
    assert out['features_length'].tolist() == [F.num_frames(int(np.ceil(10 * n / 9))) for n in (16000, 12345, 9000, 5200)]
    assert np.abs(out['features'].cpu().numpy() - oout['features']).max() < 2e-3
    # offline speeds (dataset-level speed list) through item[3], online rate 0
    batch2 = [('a', gold['pcm0'], [1], 1.1), ('b', gold['pcm1'], [2], 1.0), ('c', gold['pcm3'], [3], 0.9)]
    keys2, out2 = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, normalization=False)(batch2)
    obatch2 = [(k, (w.astype(np.float32), 16000), l, s) for k, w, l, s in batch2]
    okeys2, oout2 = K.AudioCollate(feature_extraction_conf=CONF, normalization=False)(obatch2)
    assert keys2 == okeys2 and np.abs(out2['features'].cpu().numpy() - oout2['features']).max() < 2e-3


def test_resample_rate_path(gold):
    """An 8 kHz file with resample_rate 16000 goes through the same sinc kernel as
    torchaudio.transforms.Resample (dataset.py:77-84)."""
    from openeat_b200.dataset import audio_collate_func
    x8 = signals.make('speech', 6000, 77)
    batch = [('k8', (x8, 8000), [1], 1.0), ('k16', (gold['pcm1'], 16000), [2], 1.0)]
    keys, out = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, normalization=False)(batch)
    okeys, oout = K.AudioCollate(feature_extraction_conf=CONF, normalization=False)(
        [(k, (w[0].astype(np.float32), w[1]), l, s) for k, w, l, s in batch])
    got, ref = out['features'].cpu().numpy(), oout['features']
    assert keys == okeys and got.shape == ref.shape
    # Mel bins below the old Nyquist (4 kHz ~ bin 56) carry signal: usual tolerance.  Above it the upsampled
    # audio holds only filter leakage (~-120 dB): log-mel there is ill-conditioned -- rounding the sinc table
    # to fp32 in two different ways already moves it by 4e-3 on the CPU -- so it gets a loose bound.
    assert np.abs(got - ref)[..., :56].max() < 2e-3
    assert np.abs(got - ref).max() < 5e-2


def test_fused_global_cmvn_and_stats(gold, golden_dir):
    from openeat_b200.cmvn import GlobalCMVN, load_cmvn
    from openeat_b200.dataset import audio_collate_func
    mean, istd = load_cmvn(os.path.join(golden_dir, 'cmvn_stats.json'), True)
    mean_t, istd_t = torch.from_numpy(mean).float().cuda(), torch.from_numpy(istd).float().cuda()
    batch = [('u%d' % i, gold['pcm%d' % i], [1], 1.0) for i in (0, 1, 2, 3, 5)]
    plain = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, normalization=False)
    stats = torch.zeros(161, dtype=torch.float64, device='cuda')
    fused = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, normalization=False,
                               global_cmvn=(mean_t, istd_t), cmvn_stats=stats)
    _, a = plain(batch)
    _, b = fused(batch)
    ref = GlobalCMVN(mean_t, istd_t)(a['features'])                # what the encoder does to the padded batch
    assert torch.equal(ref, b['features'])
    assert stats[160].item() == int(a['features_length'].sum())


def test_prefetching_collator_equals_direct_calls(gold):
    """H2D of batch i+1 overlapped with batch i on a side stream must not change any result."""
    from openeat_b200.dataset import PrefetchingCollator, audio_collate_func
    from openeat_b200.frontend import pack_waveforms
    conf = dict(CONF, speed_perturb_rate=0.5)
    kw = dict(data_type='wav', feature_extraction_conf=conf, normalization=True, spec_aug=True,
              spec_aug_conf=dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10))
    batches = []
    for r in range(4):
        ids = [(i + r) % 6 for i in range(5)]
        buf, offs, lens = pack_waveforms([gold['pcm%d' % i] for i in ids])
        batches.append((buf, offs, lens, ['u%d' % i for i in ids], [[i + 1] * 3 for i in ids], [1.0, 0.9, 1.1, 1.0, 1.0]))
    fn = audio_collate_func(**kw)
    random.seed(31)
    direct = [fn.collate_packed(*b) for b in batches]
    random.seed(31)
    piped = list(PrefetchingCollator(audio_collate_func(**kw), batches))
    assert len(piped) == 4
    for (k1, o1), (k2, o2) in zip(direct, piped):
        assert k1 == k2
        for name in o1:
            assert torch.equal(o1[name], o2[name]), name


@pytest.mark.parametrize('with_cmvn', [False, True])
def test_prefetching_collator_to_host_pads_on_the_host(gold, golden_dir, with_cmvn):
    """to_host=True (the reference boundary: CPU tensors, dataset.py:232-238): only the real rows cross PCIe and the
    padded tensor is completed on the host (oe_host_pad_rows) -- must equal the padded tensor the GPU path produces,
    bit for bit, padding rows included (0, or (0 - mean) * istd with a fused GlobalCMVN)."""
    from openeat_b200.cmvn import load_cmvn
    from openeat_b200.dataset import PrefetchingCollator, audio_collate_func
    from openeat_b200.frontend import pack_waveforms
    conf = dict(CONF, speed_perturb_rate=0.5)
    kw = dict(data_type='wav', feature_extraction_conf=conf, normalization=True, spec_aug=True,
              spec_aug_conf=dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10))
    if with_cmvn:
        mean, istd = load_cmvn(os.path.join(golden_dir, 'cmvn_stats.json'), True)
        kw['global_cmvn'] = (torch.from_numpy(mean).float().cuda(), torch.from_numpy(istd).float().cuda())
    batches = []
    for r in range(5):
        ids = [(i + r) % 6 for i in range(5)]
        buf, offs, lens = pack_waveforms([gold['pcm%d' % i] for i in ids])
        batches.append((buf, offs, lens, ['u%d' % i for i in ids], [[i + 1] * 3 for i in ids], [1.0, 0.9, 1.1, 1.0, 1.0]))
    fn = audio_collate_func(**kw)
    random.seed(77)
    direct = [fn.collate_packed(*b) for b in batches]
    for host_pad in (True, False):
        random.seed(77)
        got = []
        for keys, out in PrefetchingCollator(audio_collate_func(**kw), batches, to_host=True, host_pad=host_pad):
            got.append((keys, {k: v.clone() for k, v in out.items()}))     # ring slots are reused
        assert len(got) == 5
        for (k1, o1), (k2, o2) in zip(direct, got):
            assert k1 == k2
            for name in o1:
                assert not o2[name].is_cuda and o2[name].shape == o1[name].shape, name
                assert torch.equal(o1[name].cpu(), o2[name]), (name, host_pad)


def test_speed_processors(golden_dir):
    from openeat_b200.audio_processor import _speed_generator, _speed_perturb
    g = np.load(os.path.join(golden_dir, 'speed.npz'))
    a = np.load(os.path.join(golden_dir, 'augment.npz'))
    random.seed(5)
    draws = [_speed_generator([0.9, 1.1, 0.1]) for _ in range(8)] + [_speed_generator(None)] + [_speed_generator([1.05])]
    assert np.array_equal(np.array(draws), a['speed_draws'])
    random.seed(6)
    assert np.array_equal(np.array([_speed_generator([0.9, 1.1, 0]) for _ in range(8)]), a['speed_draws_uniform'])
    w = torch.from_numpy(g['x'])[None]
    assert _speed_perturb(w, 16000, 1.0) is w                      # audio_processor.py:31
    for s, tag in [(0.9, '090'), (1.1, '110')]:
        y = _speed_perturb(w, 16000, s)
        assert y.shape == (1, g['y' + tag].shape[0]) and not y.is_cuda
        assert np.abs(y[0].numpy() - g['y' + tag]).max() < 0.05
        assert _speed_perturb(w.cuda(), 16000, s).is_cuda


@pytest.mark.parametrize('i', range(5))
def test_feature_processors_bit_exact(golden_dir, i):
    from openeat_b200.feature_processor import _normalization, _spec_augmentation, _spec_substitute
    g = np.load(os.path.join(golden_dir, 'augment.npz'))
    x = g['x%d' % i]
    x0 = x.copy()
    random.seed(1000 + i)
    assert np.array_equal(_spec_augmentation(x, num_t_mask=3, num_f_mask=2, max_t=50, max_f=10), g['aug%d' % i])
    random.seed(2000 + i)
    assert np.array_equal(_spec_substitute(x, max_t=30, num_t_sub=3), g['sub%d' % i])
    random.seed(3000 + i)
    y = _spec_augmentation(_spec_substitute(x, max_t=30, num_t_sub=3), num_t_mask=3, num_f_mask=2, max_t=50, max_f=10)
    assert np.array_equal(y, g['subaug%d' % i])
    assert np.array_equal(x, x0)                                   # input untouched, new array returned
    if x.shape[0] > 1:
        assert np.abs(_normalization(x) - g['norm%d' % i]).max() < 1e-5


def test_global_cmvn_module(golden_dir):
    from openeat_b200.cmvn import GlobalCMVN, load_cmvn
    g = np.load(os.path.join(golden_dir, 'cmvn.npz'))
    mean, istd = load_cmvn(os.path.join(golden_dir, 'cmvn_stats.json'), True)
    m = GlobalCMVN(torch.from_numpy(mean).float(), torch.from_numpy(istd).float()).cuda()
    assert sorted(m.state_dict()) == ['istd', 'mean']              # checkpoint compatibility
    assert np.array_equal(m(torch.from_numpy(g['x']).cuda()).cpu().numpy(), g['y'])
    m2 = GlobalCMVN(torch.from_numpy(mean).float(), torch.from_numpy(istd).float(), norm_var=False).cuda()
    assert np.array_equal(m2(torch.from_numpy(g['x']).cuda()).cpu().numpy(), g['y_novar'])
    with pytest.raises(RuntimeError):
        m(torch.from_numpy(g['x']))                                # CPU tensor: no fallback


def test_compute_cmvn_stats_roundtrip(gold, tmp_path, tables):
    from openeat_b200.cmvn import compute_cmvn_stats, load_cmvn
    waves = [gold['pcm%d' % i] for i in (0, 1, 2, 3, 4, 5)]
    path = str(tmp_path / 'global_cmvn.json')
    s, q, n = compute_cmvn_stats([waves[:3], waves[3:]], out_json=path)
    feats = np.concatenate([F.fbank(w.astype(np.float32), window=tables[0], mel=tables[1]) for w in waves if len(w) >= 400])
    assert n == feats.shape[0]
    mean, istd = load_cmvn(path, True)                             # the reference-format consumer reads it back
    np.testing.assert_allclose(mean, feats.astype(np.float64).mean(0), rtol=1e-4)
    np.testing.assert_allclose(istd, 1.0 / feats.astype(np.float64).std(0), rtol=1e-4)


def test_wenet_style_processor_chain(gold, tmp_path, tables):
    """speed_perturb -> compute_fbank -> utt_normalize -> spec_sub -> spec_aug -> global_cmvn -> batch -> padding over
    sample dicts, against the CPU oracle with the same random draws."""
    from openeat_b200 import processor as P
    from oracle import augment as A, cmvn as C, speed as S
    lens = [16000, 9000, 5200, 12345, 300, 7777]
    lines = []
    for i in range(len(lens)):
        p = tmp_path / ('p%d.wav' % i)
        write_wav(p, gold['pcm%d' % i])
        lines.append('{"key": "k%d", "wav": "%s", "txt": "x"}' % (i, p))
    lines.append('utt:seg\tfeat:%s,0.25,0.75\tfeat_shape:0.5\ttext:x' % (tmp_path / 'p0.wav'))
    lines.append('{"key": "missing", "wav": "/nonexistent.wav", "txt": ""}')
    samples = list(P.parse_raw(lines))
    assert [s['key'] for s in samples] == ['k0', 'k1', 'k2', 'k3', 'k4', 'k5', 'seg']
    mean = torch.linspace(8.0, 12.0, 80)
    istd = torch.linspace(0.4, 0.6, 80)
    random.seed(77)
    chain = P.padding(P.batch(P.global_cmvn(P.spec_aug(P.spec_sub(P.utt_normalize(P.compute_fbank(
        P.speed_perturb(iter(samples)), num_mel_bins=80, batch_size=4)), max_t=30, num_t_sub=3),
        num_t_mask=3, num_f_mask=2, max_t=50, max_f=10), mean, istd), batch_size=16))
    keys, feats, labels, flen, llen = next(chain)
    # the oracle replays the lazy generators' draw order: every feature operation works on groups of 64 samples, so the
    # first one pulls the whole list through compute_fbank / speed_perturb (7 speed draws, in order), then every
    # surviving sample gets its spec_sub draws (in order), then every one its spec_aug draws
    random.seed(77)
    pert = [S.speed_perturb(smp['wav'].astype(np.float32), 16000, random.choice([0.9, 1.0, 1.1])) for smp in samples]
    alive = [(smp, w) for smp, w in zip(samples, pert) if len(w) >= 400]
    xs = [A.normalization(F.fbank(np.asarray(w, np.float32), window=tables[0], mel=tables[1])) for _, w in alive]
    xs = [A.spec_substitute(x, max_t=30, num_t_sub=3) for x in xs]
    xs = [A.spec_augmentation(x, 3, 2, 50, 10) for x in xs]
    out = {smp['key']: C.global_cmvn(x, mean.numpy(), istd.numpy()) for (smp, _), x in zip(alive, xs)}
    assert 'k4' not in out and set(keys) == set(out)
    assert flen.tolist() == sorted(flen.tolist(), reverse=True)
    feats = feats.cpu().numpy()
    for i, k in enumerate(keys):
        t = int(flen[i])
        assert out[k].shape[0] == t
        assert np.abs(feats[i, :t] - out[k]).max() < 3e-3
        assert np.all(feats[i, t:] == 0)


def test_kaldi_feature_path_matches_reference_chain(tmp_path, tables):
    """data_type='kaldi' (dataset.py:120-152, 190-238): features read from a binary Kaldi archive, then
    _normalization -> _spec_augmentation -> padding, against the reference's own functions (oracle mirrors of
    feature_processor.py) on the same archive with the same `random` seed.  Includes the reference's doubled
    label list (dataset.py:141,143)."""
    import random
    from openeat_b200 import kaldi_io
    from openeat_b200.dataset import audio_collate_func, _load_feature
    from oracle import augment as A
    rng = np.random.default_rng(5)
    T = [57, 120, 33, 240, 5]
    mats = {'k%d' % i: rng.normal(3.0, 2.0, (t, 80)).astype(np.float32) for i, t in enumerate(T)}
    scp = kaldi_io.write_mat_ark(str(tmp_path / 'f.ark'), mats.items())
    batch = [(k, scp[k], [i + 1, i + 2], 1.0) for i, k in enumerate(mats)]
    aug = dict(num_t_mask=2, num_f_mask=2, max_t=20, max_f=8)
    coll = audio_collate_func(data_type='kaldi', spec_aug=True, spec_aug_conf=aug, normalization=True)
    random.seed(11)
    keys, out = coll(batch)
    # the reference chain on the CPU
    random.seed(11)
    rkeys, xs, ys = _load_feature(batch)
    assert keys == rkeys == [k for k, _ in sorted(mats.items(), key=lambda kv: -kv[1].shape[0])]
    ref = [A.normalization(x) for x in xs]
    ref = [A.spec_augmentation(x, **aug) for x in ref]
    feats = out['features'].cpu().numpy()
    assert out['features_length'].tolist() == [x.shape[0] for x in xs]
    for i, r in enumerate(ref):
        got = feats[i, :r.shape[0]]
        assert np.array_equal(got == 0, r == 0)
        assert np.abs(got - r).max() <= 2e-5
        assert np.all(feats[i, r.shape[0]:] == 0)
    # doubled labels: sorted_labels[j] = labels[order[j]] over [l0, l0, l1, l1, ...]
    order = np.argsort([m.shape[0] for m in mats.values()])[::-1]
    doubled = [lab for i in range(len(T)) for lab in ([i + 1, i + 2], [i + 1, i + 2])]
    assert out['targets'].cpu().numpy()[:, :2].tolist() == [doubled[j] for j in order]


def test_feature_dither_statistics_and_masks():
    """feature_dither (dataset.py:199-201): x + (U[0,1) - 0.5) * a with a = random.uniform(0, fd) drawn like the
    reference; stochastic, so only the distribution is checked: bounded by a/2, mean 0, variance a^2/12; masked
    cells stay exactly 0; the `random` stream is consumed exactly as the reference does (speed gates, the dither
    draw, then the mask draws); two batches get different noise."""
    import random
    from openeat_b200 import planner
    from openeat_b200.dataset import audio_collate_func
    from oracle import signals
    lens = [48000, 32000, 16000]
    frames = np.array([298, 198, 98], np.int32)
    items = [('u%d' % i, signals.make('speech', n, 20 + i), [1], 1.0) for i, n in enumerate(lens)]
    conf = {'mel_bins': 80, 'speed_perturb_rate': 0, 'speeds': [1.0], 'wav_dither': 0.0}
    aug = dict(num_t_mask=1, num_f_mask=1, max_t=10, max_f=5)
    kw = dict(data_type='wav', feature_extraction_conf=conf, normalization=True, spec_aug=True, spec_aug_conf=aug)
    plain = audio_collate_func(**kw)
    dith = audio_collate_func(feature_dither=0.5, **kw)
    # what the reference does with `random` for this batch: 3 speed gates, the dither amplitude, then the masks
    random.seed(3)
    for _ in items:
        random.random()
    a = random.uniform(0, 0.5)
    _, tm, fm = planner.plan_augment(frames, 80, None, aug)
    expect_next = random.random()
    random.seed(3)
    _, o1 = dith(items)
    assert random.random() == expect_next                      # same number of draws, same order
    random.seed(3)
    _, o2 = dith(items)
    _, ob = plain(items)                                       # other masks (its stream has no dither draw)
    y1, y2, yb = (o['features'].cpu().numpy() for o in (o1, o2, ob))
    assert not np.array_equal(y1, y2)                          # a new Philox key per batch
    noise = []
    for i, t in enumerate(frames):
        keep = np.ones((t, 80), bool)
        for s_, e_ in tm[i]:
            keep[s_:e_] = False
        for s_, e_ in fm[i]:
            keep[:, s_:e_] = False
        assert np.all(y1[i, :t][~keep] == 0) and np.all(y2[i, :t][~keep] == 0)
        both = keep & (yb[i, :t] != 0)
        noise.append((y1[i, :t] - yb[i, :t])[both])
        assert np.all(y1[i, t:] == 0)
    noise = np.concatenate(noise)
    assert noise.size > 20000
    assert np.abs(noise).max() <= a / 2 + 1e-5
    assert abs(noise.mean()) < 5 * a / np.sqrt(12 * noise.size) + 1e-6
    assert abs(noise.var() - a * a / 12) < 0.05 * a * a / 12


def test_wav_dither_through_the_collate_mirror():
    """feature_extraction_conf['wav_dither'] (dataset.py:98) through audio_collate_func: with dither the speed perturb is
    not fused (the library refuses that combination), the mirror resamples first and dithers the fp32 result; frame
    counts, ordering and the consumption of Python's `random` stream are those of the undithered run."""
    import random
    from openeat_b200.dataset import audio_collate_func
    from oracle import signals
    lens = [30000, 48000, 20000, 16000]
    items = [('u%d' % i, signals.make('speech', n, 60 + i), [1, 2], s) for i, (n, s) in enumerate(zip(lens, (0.9, 1.0, 1.1, 1.0)))]
    base = {'mel_bins': 80, 'speed_perturb_rate': 0, 'speeds': [1.0], 'wav_dither': 0.0}
    outs = []
    for d in (0.0, 1.0):
        coll = audio_collate_func(data_type='wav', feature_extraction_conf=dict(base, wav_dither=d), normalization=False)
        random.seed(5)
        keys, out = coll(items)
        outs.append((keys, out, random.random()))
    (k0, o0, r0), (k1, o1, r1) = outs
    assert k0 == k1 and r0 == r1
    assert o0['features_length'].tolist() == o1['features_length'].tolist()
    a, b = o0['features'].cpu().numpy(), o1['features'].cpu().numpy()
    assert np.isfinite(b).all() and not np.array_equal(a, b)
    assert np.median(np.abs(a - b)) < 0.05                         # speech at +/- 3000 vs dither 1.0


@pytest.mark.parametrize('tag,kw', [
    ('plain', dict(normalization=False)),
    ('norm_aug', dict(normalization=True, spec_aug=True, spec_aug_conf=dict(num_t_mask=2, num_f_mask=2, max_t=20, max_f=8))),
    ('sub', dict(normalization=True, spec_sub=True, spec_sub_conf=dict(num_t_sub=3, max_t=10))),
])
def test_kaldi_collate_matches_reference_output(tmp_path, golden_dir, tag, kw):
    """The REFERENCE's audio_collate_func(data_type='kaldi') run end to end under oracle/ref_shim.py
    (tests/golden/kaldi_collate.npz, oracle/make_golden.py section 8) against this package's collate on the same
    archive and `random` seed: key order (ties included), the doubled-label quirk of dataset.py:141-143, lengths,
    padding, masks bit-exact, values within fp32 normalisation noise."""
    import random
    from openeat_b200 import kaldi_io
    from openeat_b200.dataset import audio_collate_func
    g = np.load(os.path.join(golden_dir, 'kaldi_collate.npz'))
    n = len(g['labels'])
    mats = [g['mat%d' % i] for i in range(n)]
    scp = kaldi_io.write_mat_ark(str(tmp_path / 'feats.ark'), [('k%d' % i, m) for i, m in enumerate(mats)])
    batch = [('k%d' % i, scp['k%d' % i], [i + 1] * int(g['labels'][i]), 1.0) for i in range(n)]
    coll = audio_collate_func(data_type='kaldi', **kw)
    random.seed(777)
    keys, out = coll([batch])
    assert keys == g[tag + '_keys'].tolist()
    assert out['features_length'].tolist() == g[tag + '_features_length'].tolist()
    assert out['targets'].cpu().numpy().tolist() == g[tag + '_targets'].tolist()
    assert out['targets_length'].tolist() == g[tag + '_targets_length'].tolist()
    got, ref = out['features'].cpu().numpy(), g[tag + '_features']
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.array_equal(got == 0, ref == 0)
    assert np.abs(got - ref).max() <= (0 if tag == 'plain' else 2e-5)


def test_long_ratio_and_kaiser_resamplers(golden_dir):
    """Ratios whose polyphase table is too long to keep (44.1 kHz -> 16 kHz = 441:160; a speed drawn from a continuous
    range, 953:1000) are evaluated on the fly (OE_RS_DIRECT) -- against torchaudio.functional.resample goldens
    (oracle/make_golden_r02.py).  A caller-designed table with an arbitrary tap count (the sox-quality Kaiser filter)
    registers and runs through the generic kernel -- against the oracle's own design of the same specification."""
    from openeat_b200.frontend import Frontend, kaiser_sinc_kernel, pack_waveforms
    from oracle import speed as S
    fe = Frontend(mel_bins=80, sample_rate=16000)
    g = np.load(os.path.join(golden_dir, 'resample_long.npz'))
    for name, (o, n), tol in (('441_160', (441, 160), 0.1), ('953_1000', (953, 1000), 1.0)):
        x, ref = g['x_' + name], g['y_' + name]
        exact = S.resample(x, o, n, dtype=np.float64)                   # the same filter, evaluated in float64
        for dtype in (np.float32, np.int16):
            buf, offs, lens = pack_waveforms([x.astype(dtype)], dtype=dtype)
            out, ooffs, olens = fe.resample(buf.cuda(), offs, lens, [(o, n)])
            torch.cuda.synchronize()
            got = out[ooffs[0]:ooffs[0] + int(olens[0])].cpu().numpy()
            assert got.shape == ref.shape
            assert np.abs(got - exact).max() <= 0.02                    # int16 scale
            assert np.abs(got - ref).max() <= tol                       # torchaudio's fp32 kernel grid is that noisy (see the oracle test)
    # the whole collate path with a 44.1 kHz source: resample_rate stage through the on-the-fly resampler
    from openeat_b200.dataset import _extract_feature
    from oracle import fbank as F
    keys, feats, _ = _extract_feature([('u', (g['x_441_160'].astype(np.int16), 44100), [1], 1.0)],
                                      {'mel_bins': 80, 'resample_rate': 16000, 'speed_perturb_rate': 0, 'wav_dither': 0.0})
    ref = F.fbank(g['y_441_160'])
    assert keys == ['u'] and feats[0].shape == ref.shape and np.abs(feats[0][:, :72] - ref[:, :72]).max() <= 5e-3
    # Kaiser design: product vs oracle table, then the resampled waveform vs the oracle's float64 resampling
    for sp in (0.9, 1.1):
        o, n = S.speed_ratio(sp)
        k, width = kaiser_sinc_kernel(o, n)
        kr, wr = S.soxlike_kernel(o, n, dtype=np.float64)
        assert width == wr and np.abs(k - kr).max() <= 1e-7
        x = g['x_953_1000']
        buf, offs, lens = pack_waveforms([x], dtype=np.float32)
        out, ooffs, olens = fe.resample(buf.cuda(), offs, lens, [(o, n)], kind='kaiser')
        torch.cuda.synchronize()
        got = out[ooffs[0]:ooffs[0] + int(olens[0])].cpu().numpy()
        ref = S.speed_perturb_soxlike(x, 16000, sp)
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 0.05


def test_cmvn_conv_subsample_matches_torch():
    """GlobalCMVN + Conv2d(1, odim, 3, 2) + ReLU in one kernel against the fp32 torch reference of the same
    two operations (encoder.py:221-223, subsampling.py:76-78, 110-111) -- odd and even T, odim 256 and 32, with and
    without CMVN / variance normalisation / bias, tolerance 1e-4 (nine-term fp32 dot products)."""
    from openeat_b200.cmvn import GlobalCMVN
    from openeat_b200.subsampling import CmvnConvSubsample
    torch.manual_seed(3)
    for B, T, F, odim, use_cmvn, norm_var, bias in ((3, 67, 80, 256, True, True, True), (2, 16, 80, 32, True, False, True),
                                                  (1, 3, 80, 8, False, True, False), (2, 101, 40, 64, True, True, True)):
        x = (torch.randn(B, T, F) * 4.0 + 10.0).cuda()
        conv = torch.nn.Conv2d(1, odim, 3, 2, bias=bias).cuda()
        cm = GlobalCMVN(torch.randn(F).cuda() + 10.0, torch.rand(F).cuda() + 0.2, norm_var=norm_var) if use_cmvn else None
        got = CmvnConvSubsample(conv, cm)(x)
        with torch.no_grad():
            z = x
            if cm is not None:
                z = z - cm.mean
                if norm_var:
                    z = z * cm.istd
            prev = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = False
            ref = torch.relu(conv(z.unsqueeze(1)))
            torch.backends.cudnn.allow_tf32 = prev
        assert got.shape == ref.shape == (B, odim, (T - 3) // 2 + 1, (F - 3) // 2 + 1)
        assert float((got - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))


def test_collate_reads_24bit_and_float_wav(gold, tmp_path):
    """torchaudio.load reads more than 16-bit PCM (dataset.py:62-75); a batch mixing 24-bit PCM, IEEE-float and 16-bit
    files goes through the fp32 waveform path and matches the oracle's fbank of the decoded values."""
    import struct
    from openeat_b200.dataset import audio_collate_func, read_wav

    def riff(path, tag, bits, raw, sr=16000):
        fmt = struct.pack('<HHIIHH', tag, 1, sr, sr * bits // 8, bits // 8, bits)
        body = b'WAVE' + b'fmt ' + struct.pack('<I', len(fmt)) + fmt + b'data' + struct.pack('<I', len(raw)) + raw
        with open(path, 'wb') as f:
            f.write(b'RIFF' + struct.pack('<I', len(body)) + body)

    pcm0, pcm1, pcm3 = gold['pcm0'], gold['pcm1'], gold['pcm3']
    v24 = pcm0.astype(np.int64) * 256 + 77                                   # 24-bit samples that are not multiples of 256
    raw24 = np.stack([v24 & 255, (v24 >> 8) & 255, (v24 >> 16) & 255], axis=-1).astype(np.uint8).tobytes()
    riff(str(tmp_path / 'a24.wav'), 1, 24, raw24)
    riff(str(tmp_path / 'bf.wav'), 3, 32, (pcm1.astype(np.float32) / 32768 * 0.5).astype('<f4').tobytes())
    write_wav(tmp_path / 'c16.wav', pcm3)
    batch = [('a', str(tmp_path / 'a24.wav'), [1], 1.0), ('b', str(tmp_path / 'bf.wav'), [2], 1.0),
             ('c', str(tmp_path / 'c16.wav'), [3], 1.0)]
    keys, out = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, normalization=False)(batch)
    assert keys == ['a', 'c', 'b']                                           # 16000, 12345, 9000 samples
    feats = out['features'].cpu().numpy()
    for row, name in enumerate(('a24.wav', 'c16.wav', 'bf.wav')):
        x, sr = read_wav(str(tmp_path / name))
        ref = F.fbank(np.asarray(x, dtype=np.float32))
        assert np.abs(feats[row, :ref.shape[0]] - ref).max() < 1e-3
    assert np.abs(read_wav(str(tmp_path / 'a24.wav'))[0] - (pcm0.astype(np.float32) + np.float32(77 / 256))).max() < 1e-2


@pytest.mark.parametrize('seed', range(12))
def test_random_batches_against_the_oracle_collate(seed):
    """Randomised end-to-end parity: random batch sizes, lengths around the framing edges (399 / 400 / 559 / 560 samples,
    multiples of the tile), signal classes, offline speeds, normalisation / spec_sub / spec_aug switches -- the GPU collate
    against the oracle's port of audio_collate_func with the same Python-random seed: same keys (order, drops), same
    lengths, same mask / substitution pattern bit for bit, values within the stated tolerances."""
    from openeat_b200.dataset import audio_collate_func
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 24))
    edge = [399, 400, 401, 559, 560, 561, 5360, 5361, 5520, 400 + 160 * 31, 400 + 160 * 32]
    lens = [int(rng.choice(edge)) if rng.random() < 0.4 else int(rng.integers(300, 40000)) for _ in range(B)]
    norm = bool(rng.integers(2))
    # per-utterance normalisation of (near-)constant features is 0 / 0 in the reference: the
    # degenerate signal classes only take part without it
    kinds = ['white', 'speech'] if norm else ['white', 'speech', 'lsb', 'zero', 'square']
    waves = [signals.make(kinds[int(rng.integers(len(kinds)))], n, 50 * seed + i) for i, n in enumerate(lens)]
    speeds = [float(rng.choice([0.9, 1.0, 1.0, 1.1])) for _ in range(B)]
    kw = dict(normalization=norm)
    if rng.integers(2):
        kw.update(spec_aug=True, spec_aug_conf=dict(num_t_mask=int(rng.integers(1, 4)), num_f_mask=int(rng.integers(1, 3)),
                                                    max_t=int(rng.integers(5, 60)), max_f=int(rng.integers(2, 12))))
    if rng.integers(2):
        kw.update(spec_sub=True, spec_sub_conf=dict(num_t_sub=int(rng.integers(1, 4)), max_t=int(rng.integers(5, 40))))
    batch = [('k%d' % i, w, list(range(1, 2 + i % 5)), s) for i, (w, s) in enumerate(zip(waves, speeds))]
    obatch = [(k, (w.astype(np.float32), 16000), l, s) for k, w, l, s in batch]
    random.seed(seed)
    keys, out = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, **kw)(batch)
    random.seed(seed)
    okeys, oout = K.AudioCollate(feature_extraction_conf=CONF, **kw)(obatch)
    assert list(keys) == list(okeys)
    if not okeys:
        assert out['features'].numel() == 0
        return
    assert np.array_equal(out['features_length'].cpu().numpy(), oout['features_length'])
    assert np.array_equal(out['targets'].cpu().numpy(), oout['targets'])
    got, ref = out['features'].cpu().numpy(), oout['features']
    assert got.shape == ref.shape
    ok = np.ones(len(okeys), bool)
    if norm:                                                            # one or two frames: std = 0 or a single rounding
        ok &= oout['features_length'] >= 3
    zero_ref = (ref == 0)
    assert np.array_equal((got == 0)[ok], zero_ref[ok])                 # masks, substitutions, padding: bit-exact pattern
    tol = 5e-3 if norm else 2e-3                                        # normalisation divides by per-bin std (>= 1e-3 here)
    d = np.abs(got - ref)[ok]
    assert np.nanmax(d) < tol * max(1.0, float(np.abs(ref[ok]).max()) / 10.0), float(np.nanmax(d))
