"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU oracle and the golden vectors.

Tolerances (BASELINE.json north_star):
  log-mel  max-abs <= 1e-3 vs the fp32 oracle/golden on noise- and speech-like signals, and
           <= max(1e-3, fp32-vs-fp64 oracle gap) vs the fp64 golden on tonal signals;
  masks / substitution indices / padding: bit-exact; run-to-run: bit-identical.
"""
import os
import random

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import augment as A      # noqa: E402
from oracle import cmvn as C         # noqa: E402
from oracle import fbank as F        # noqa: E402
from oracle import signals           # noqa: E402

CASES = ['fbank_white_400', 'fbank_white_559', 'fbank_white_560', 'fbank_white_8000', 'fbank_white_80000',
         'fbank_white_560000', 'fbank_speech_8000', 'fbank_speech_80000', 'fbank_lsb_8000', 'fbank_zero_8000',
         'fbank_dcsine_8000', 'fbank_square_8000']


@pytest.fixture(scope='module')
def fe():
    from openeat_b200.frontend import Frontend
    return Frontend(mel_bins=80, sample_rate=16000)


def run_raw(fe, waves, dtype=np.int16, **kw):
    from openeat_b200.frontend import pack_waveforms
    buf, offs, lens = pack_waveforms([np.asarray(w, dtype=dtype) for w in waves], dtype=dtype)
    out, frames = fe.fbank(buf.cuda(), offs, lens, **kw)
    torch.cuda.synchronize()
    return (out.cpu().numpy() if out is not None else None), frames


def test_tables_are_torchaudios(fe, tables):
    win, mel = fe.tables()
    assert np.array_equal(win, tables[0]) and np.array_equal(mel, tables[1])


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('dtype', [np.int16, np.float32])
def test_fbank_matches_golden(fe, name, dtype, golden_dir, manifest):
    meta = manifest[name]
    x = signals.make(meta['kind'], meta['samples'], meta['seed'])
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    stride = int(g['stride'])
    y, frames = run_raw(fe, [x], dtype=dtype, layout='ragged')
    assert frames.tolist() == [meta['frames']] and y.shape == (meta['frames'], 80)
    assert np.isfinite(y).all()
    tol = max(1e-3, meta['gap32_64'])
    assert np.abs(y[::stride] - g['y64']).max() <= tol, 'vs fp64 torchaudio'
    if meta['kind'] in ('white', 'speech', 'lsb', 'zero', 'square'):
        assert np.abs(y[::stride] - g['y32']).max() <= 1e-3, 'vs fp32 torchaudio'
    if meta['kind'] in ('white', 'speech', 'lsb'):     # random signals: errors must average out (no systematic bias);
        # periodic signals repeat the same rounding pattern in every frame, so nothing averages there
        np.testing.assert_allclose(y.astype(np.float64).sum(0), g['colsum64'], rtol=0, atol=0.25 * tol * meta['frames'])


def test_zero_signal_is_exactly_log_eps(fe):
    y, _ = run_raw(fe, [np.zeros(8000, np.int16)], layout='ragged')
    assert np.abs(y - np.log(F.EPS_F32)).max() < 2e-6


def test_ragged_batch_padded_layout_and_drop(fe, tables):
    """Mixed lengths in one launch, incl. < 400 samples (0 frames: the reference drops it),
    exactly one frame, tile-boundary lengths; padding must be exactly 0."""
    lens = [399, 400, 559, 560, 160 * 31 + 400, 160 * 32 + 400, 160 * 33 + 399, 12345, 0, 80000]
    waves = [signals.make('speech' if i % 2 else 'white', n, 100 + i) for i, n in enumerate(lens)]
    y, frames = run_raw(fe, waves, layout='padded')
    exp = [F.num_frames(n) for n in lens]
    assert frames.tolist() == exp
    assert y.shape == (len(lens), max(exp), 80)
    for i, w in enumerate(waves):
        if exp[i]:
            ref = F.fbank(w.astype(np.float32), window=tables[0], mel=tables[1])
            assert np.abs(y[i, :exp[i]] - ref).max() <= 1e-3
        assert np.all(y[i, exp[i]:] == 0.0)


def test_run_to_run_bitwise_stable(fe):
    waves = [signals.make('speech', n, 7 + i) for i, n in enumerate([30000, 5000, 77777])]
    a, _ = run_raw(fe, waves, layout='padded', normalization=True)
    for _ in range(3):
        b, _ = run_raw(fe, waves, layout='padded', normalization=True)
        assert np.array_equal(a, b, equal_nan=True)


def test_frames_independent_streaming_windows(fe):
    """BASELINE config 5: a stream cut into 16-frame windows (2800 samples, 240 overlap)
    gives exactly the whole-stream features."""
    x = signals.make('speech', 160 * 16 * 40 + 240, 5)
    whole, fr = run_raw(fe, [x], layout='ragged')
    wins = [x[2560 * i: 2560 * i + 2800] for i in range(40)]
    parts, frames = run_raw(fe, wins, layout='ragged')
    assert frames.tolist() == [16] * 40 and fr[0] == 640
    assert np.array_equal(parts, whole)


def _plans(frames, seed, n_t=3, n_f=2, max_t=50, max_f=10, sub=False):
    random.seed(seed)
    subs = [A.plan_spec_substitute(int(t), max_t=30, num_t_sub=3) for t in frames] if sub else None
    plans = [A.plan_spec_augmentation(int(t), 80, n_t, n_f, max_t, max_f) for t in frames]
    tm = np.array([p[0] for p in plans], dtype=np.int32)
    fm = np.array([p[1] for p in plans], dtype=np.int32)
    return subs, plans, tm, fm


def test_fused_specaug_cmvn_single_pass(fe, tables, golden_dir):
    lens = [16000, 9000, 5200, 12345]
    waves = [signals.make('speech', n, 200 + i) for i, n in enumerate(lens)]
    frames = [F.num_frames(n) for n in lens]
    _, plans, tm, fm = _plans(frames, 11)
    mean, istd = C.load_cmvn(os.path.join(golden_dir, 'cmvn_stats.json'), True)
    mean_d = torch.from_numpy(mean).float().cuda()
    istd_d = torch.from_numpy(istd).float().cuda()
    for on_pad in (False, True):
        y, _ = run_raw(fe, waves, layout='padded', tmask=tm, fmask=fm, cmvn=(mean_d, istd_d),
                       cmvn_on_padding=on_pad)
        for i, w in enumerate(waves):
            raw = F.fbank(w.astype(np.float32), window=tables[0], mel=tables[1])
            aug = A.apply_spec_augmentation(raw, *plans[i])
            ref = C.global_cmvn(aug, mean, istd)
            got = y[i, :frames[i]]
            masked = aug == 0
            assert np.array_equal(got[masked], ref[masked])            # masked cells: exactly (0-mean)*istd
            assert np.abs(got - ref).max() <= 1e-3 * float(istd.max())
            pad = y[i, frames[i]:]
            if on_pad:
                assert np.array_equal(pad, np.broadcast_to(C.global_cmvn(np.zeros(80), mean, istd), pad.shape))
            else:
                assert np.all(pad == 0)


@pytest.mark.parametrize('sub', [False, True])
def test_two_phase_norm_sub_aug(fe, tables, sub):
    lens = [16000, 9000, 5200, 12345, 700, 48000]
    waves = [signals.make('speech' if i % 2 else 'white', n, 300 + i) for i, n in enumerate(lens)]
    frames = [F.num_frames(n) for n in lens]
    subs, plans, tm, fm = _plans(frames, 12, sub=sub)
    maps = [A.substitute_index_map(t, s) for t, s in zip(frames, subs)] if sub else None
    y, _ = run_raw(fe, waves, layout='padded', normalization=True, tmask=tm, fmask=fm, frame_maps=maps)
    for i, w in enumerate(waves):
        raw = F.fbank(w.astype(np.float32), window=tables[0], mel=tables[1])
        ref = A.normalization(raw)
        if sub:
            ref = A.apply_spec_substitute(ref, subs[i])
        ref = A.apply_spec_augmentation(ref, *plans[i])
        got = y[i, :frames[i]]
        assert np.array_equal(got == 0, ref == 0)                      # masks bit-exact
        assert np.abs(got - ref).max() <= 2e-3                         # 1e-3 log-mel / std(~0.5+)
        assert np.all(y[i, frames[i]:] == 0)


def test_normalization_zero_variance_is_degenerate_like_reference(fe):
    """feature_processor.py:8 has no epsilon: constant features are 0/0.  numpy itself returns NaN for
    some lengths and a +-1 rounding artefact for others (fp32 mean of n equal values is not exact), so
    there is no value parity in this regime -- only: no epsilon is added, i.e. NaN or |y| <= 1."""
    for n in (2000, 4000, 16000):
        y, _ = run_raw(fe, [np.zeros(n, np.int16)], layout='padded', normalization=True)
        assert (np.isnan(y) | (np.abs(y) <= 1.0 + 1e-6)).all()


def test_cmvn_stats_accumulate(fe, tables):
    lens = [16000, 9000, 300, 5200, 80000]
    waves = [signals.make('speech', n, 400 + i) for i, n in enumerate(lens)]
    stats = torch.zeros(161, dtype=torch.float64, device='cuda')
    _, frames = run_raw(fe, waves[:2], layout='ragged', stats=stats, want_out=False)
    y, frames2 = run_raw(fe, waves[2:], layout='ragged', stats=stats)
    feats = [F.fbank(w.astype(np.float32), window=tables[0], mel=tables[1]) for w in waves if len(w) >= 400]
    s, q, n = C.compute_cmvn_stats(feats)
    got = stats.cpu().numpy()
    assert got[160] == n
    np.testing.assert_allclose(got[:80], s, rtol=1e-4)                 # north_star: 1e-4 relative
    np.testing.assert_allclose(got[80:160], q, rtol=1e-4)
    # and against the device features themselves (tight): the reduction adds no error of its own
    s2, q2, _ = C.compute_cmvn_stats([y])
    rest = C.compute_cmvn_stats(feats[:2])
    np.testing.assert_allclose(got[:80] - rest[0], s2, rtol=1e-5)


def test_features_in_mode_matches_reference_processors(fe, golden_dir):
    """The numpy-level reference processors through the device finalize path (features in)."""
    g = np.load(os.path.join(golden_dir, 'augment.npz'))
    xs = [g['x0'], g['x1'], g['x2']]
    rows = np.array([x.shape[0] for x in xs], dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(rows[:-1])]).astype(np.int64)
    feats = torch.from_numpy(np.concatenate(xs)).cuda()
    plans, maps = [], []
    for i, x in enumerate(xs):
        random.seed(3000 + i)
        subs = A.plan_spec_substitute(x.shape[0], max_t=30, num_t_sub=3)
        maps.append(A.substitute_index_map(x.shape[0], subs))
        plans.append(A.plan_spec_augmentation(x.shape[0], 80, 3, 2, 50, 10))
    tm = np.array([p[0] for p in plans], dtype=np.int32)
    fm = np.array([p[1] for p in plans], dtype=np.int32)
    out, _ = fe.fbank(feats, offs, rows, layout='padded', features_in=True, tmask=tm, fmask=fm, frame_maps=maps)
    out = out.cpu().numpy()
    for i, x in enumerate(xs):
        assert np.array_equal(out[i, :rows[i]], g['subaug%d' % i])     # bit-exact vs the reference's own output
    out, _ = fe.fbank(feats, offs, rows, layout='padded', features_in=True, normalization=True)
    out = out.cpu().numpy()
    for i, x in enumerate(xs):
        assert np.abs(out[i, :rows[i]] - g['norm%d' % i]).max() < 1e-5


def test_speed_perturb_resampler(fe, golden_dir):
    from openeat_b200.frontend import pack_waveforms
    g = np.load(os.path.join(golden_dir, 'speed.npz'))
    x = g['x']
    buf, offs, lens = pack_waveforms([x, x, x, x.astype(np.int16)[:3000]], dtype=np.float32)
    out, ooffs, olens = fe.resample(buf.cuda(), offs, lens, [(9, 10), (11, 10), None, (9, 10)])
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    assert olens.tolist() == [8889, 7273, 8000, 3334]
    y090 = out[ooffs[0]:ooffs[0] + olens[0]]
    y110 = out[ooffs[1]:ooffs[1] + olens[1]]
    assert np.abs(y090 - g['y090']).max() < 0.05                        # int16 scale: stated resampler tolerance
    assert np.abs(y110 - g['y110']).max() < 0.05
    assert np.array_equal(out[ooffs[2]:ooffs[2] + 8000], x)
    # int16 input path + fbank of the perturbed audio vs torchaudio speed -> fbank
    buf16, o16, l16 = pack_waveforms([x.astype(np.int16)], dtype=np.int16)
    r, ro, rl = fe.resample(buf16.cuda(), o16, l16, [(9, 10)])
    y, frames = fe.fbank(r, ro, rl, layout='ragged')
    assert np.abs(y.cpu().numpy() - g['fb090']).max() < 2e-3


def test_fused_speed_perturb_equals_separate_resampler(fe, golden_dir):
    """Speed perturb fused into the fbank kernel's staging (no intermediate waveform) == oe_resample followed
    by fbank on its fp32 output -- same taps, same summation order, so bit for bit -- across tile
    boundaries, utterance ends and utterances the perturbation pushes below / above one window."""
    from openeat_b200.frontend import pack_waveforms
    lens = [16000, 9000, 50000, 5200, 7777, 400, 450, 5121 * 9 // 10, 361]
    waves = [signals.make('speech' if i % 2 else 'white', n, 600 + i) for i, n in enumerate(lens)]
    ratios = np.array([[9, 10], [11, 10], [9, 10], [0, 0], [11, 10], [9, 10], [11, 10], [9, 10], [9, 10]])
    buf, offs, ln = pack_waveforms(waves)
    dev = buf.cuda()
    fused, frames = fe.fbank(dev, offs, ln, layout='padded', speed_ratios=ratios)
    r, ro, rl = fe.resample(dev, offs, ln, ratios)
    sep, frames2 = fe.fbank(r, ro, rl, layout='padded')
    torch.cuda.synchronize()
    assert frames.tolist() == frames2.tolist()
    assert frames.tolist() == [F.num_frames(-(-n * b // a) if a else n) for n, (a, b) in zip(lens, ratios.tolist())]
    assert torch.equal(fused, sep)
    # and against torchaudio.functional.speed -> kaldi.fbank (golden)
    g = np.load(os.path.join(golden_dir, 'speed.npz'))
    b16, o16, l16 = pack_waveforms([g['x'].astype(np.int16)] * 2)
    y, _ = fe.fbank(b16.cuda(), o16, l16, layout='padded', speed_ratios=np.array([[9, 10], [11, 10]]))
    y = y.cpu().numpy()
    assert np.abs(y[0, :g['fb090'].shape[0]] - g['fb090']).max() < 2e-3
    assert np.abs(y[1, :g['fb110'].shape[0]] - g['fb110']).max() < 2e-3
    with pytest.raises(Exception):
        fe.fbank(dev, offs, ln, layout='padded', speed_ratios=np.array([[1, 2]] * len(lens)))   # not fusable: says so


def test_global_cmvn_apply(fe, golden_dir):
    g = np.load(os.path.join(golden_dir, 'cmvn.npz'))
    mean = torch.from_numpy(g['mean_json']).float().cuda()
    istd = torch.from_numpy(g['istd_json']).float().cuda()
    x = torch.from_numpy(g['x']).cuda()
    assert np.array_equal(fe.cmvn_apply(x, mean, istd).cpu().numpy(), g['y'])
    assert np.array_equal(fe.cmvn_apply(x, mean, None).cpu().numpy(), g['y_novar'])


def test_errors_are_reported_not_fatal(fe):
    from openeat_b200 import FrontendError
    buf = torch.zeros(1000, dtype=torch.int16, device='cuda')
    with pytest.raises(FrontendError):
        fe.fbank(buf, np.array([3]), np.array([500]))                   # misaligned offset
    with pytest.raises(FrontendError):
        fe.fbank(buf.double(), np.array([0]), np.array([500]))          # unsupported dtype
    from openeat_b200.frontend import Frontend
    with pytest.raises(FrontendError):
        Frontend(mel_bins=80, sample_rate=8000)                         # unsupported framing says so


def test_generation_1_and_2_kernels_agree(tables, monkeypatch):
    """The second-generation fbank kernel (radix-4 stages, pair untangle, per-group power slices) against the
    first-generation one (OE_FBANK_V1=1 at handle creation) and the oracle, int16 and fp32 input, with a fused
    speed perturb in the batch: both within 1e-3 of the oracle, and within 3e-4 of each other."""
    from openeat_b200.frontend import Frontend
    fe2 = Frontend(mel_bins=80, sample_rate=16000)
    monkeypatch.setenv('OE_FBANK_V1', '1')
    fe1 = Frontend(mel_bins=80, sample_rate=16000)
    monkeypatch.delenv('OE_FBANK_V1')
    lens = [400, 559, 5200, 16000, 33333, 80000]
    waves = [signals.make(('white', 'speech', 'lsb')[i % 3], n, 40 + i) for i, n in enumerate(lens)]
    for dtype in (np.int16, np.float32):
        y1, f1 = run_raw(fe1, waves, dtype=dtype, layout='padded')
        y2, f2 = run_raw(fe2, waves, dtype=dtype, layout='padded')
        assert f1.tolist() == f2.tolist()
        assert np.abs(y1 - y2).max() <= 3e-4
        for i, w in enumerate(waves):
            ref = F.fbank(w.astype(np.float32), window=tables[0], mel=tables[1])
            assert np.abs(y2[i, :f2[i]] - ref).max() <= 1e-3 and np.abs(y1[i, :f1[i]] - ref).max() <= 1e-3
    ratios = np.array([[0, 0], [9, 10], [11, 10], [0, 0], [9, 10], [11, 10]], np.int32)
    y1, f1 = run_raw(fe1, waves, layout='padded', speed_ratios=ratios)
    y2, f2 = run_raw(fe2, waves, layout='padded', speed_ratios=ratios)
    # speed 0.9 leaves the top mel bins with filter leakage only (ill-conditioned): the two roundings may differ more there
    assert f1.tolist() == f2.tolist() and np.abs(y1 - y2).max() <= 1e-3


def test_single_pass_padding_rows_and_launch_count(fe, tables):
    """Padded single-pass layout: the fbank kernel writes whole 32-frame tiles, oe_pad_fill_kernel the rows behind
    them.  Frame counts around the tile size (0 = dropped utterance, 1, 31, 32, 33, 64, 65) and a tensor much
    longer than every utterance; the library's own launch counter sees descriptor + fbank + fill."""
    frames = [0, 1, 31, 32, 33, 64, 65]
    lens = [300] + [400 + 160 * (m - 1) for m in frames[1:]]
    waves = [signals.make('white', n, 70 + i) for i, n in enumerate(lens)]
    rng = np.random.default_rng(3)
    mean, istd = rng.normal(10, 1, 80).astype(np.float32), rng.uniform(0.3, 0.7, 80).astype(np.float32)
    mean_d, istd_d = torch.from_numpy(mean).cuda(), torch.from_numpy(istd).cuda()
    from openeat_b200.frontend import pack_waveforms
    buf, offs, ln = pack_waveforms(waves)
    dev = buf.cuda()
    tmax = 200
    out = torch.full((len(waves), tmax, 80), float('nan'), device='cuda')
    rows = np.arange(len(waves), dtype=np.int64) * tmax
    nrows = np.full(len(waves), tmax, np.int32)
    for on_pad in (False, True):
        out.fill_(float('nan'))
        n0 = fe.launches
        _, got_frames = fe.fbank(dev, offs, ln, layout='custom', out=out.view(-1, 80), out_rows=rows, out_nrows=nrows,
                                 cmvn=(mean_d, istd_d), cmvn_on_padding=on_pad)
        assert fe.launches - n0 == 3         # descriptors (the small batch carries its metadata in the kernel parameters), fbank, padding fill
        torch.cuda.synchronize()
        y = out.cpu().numpy()
        assert got_frames.tolist() == frames and np.isfinite(y).all()
        pad_val = C.global_cmvn(np.zeros(80), mean, istd) if on_pad else np.zeros(80, np.float32)
        for i, w in enumerate(waves):
            if frames[i]:
                ref = C.global_cmvn(F.fbank(w.astype(np.float32), window=tables[0], mel=tables[1]), mean, istd)
                assert np.abs(y[i, :frames[i]] - ref).max() <= 1e-3 * float(istd.max())
            assert np.array_equal(y[i, frames[i]:], np.broadcast_to(pad_val, (tmax - frames[i], 80)))
    # the benchmark configuration (per-utterance normalisation + global statistics): descriptors, the fbank kernel (raw rows
    # straight into the padded tensor, global statistics accumulated in the kernel) and the in-place completion
    # (oe_finalize2_kernel: statistics merge + normalisation + padding); the dropped utterance (0 frames) gets its padding
    stats = torch.zeros(161, dtype=torch.float64, device='cuda')
    n0 = fe.launches
    y, fr = fe.fbank(dev, offs, ln, layout='padded', normalization=True, stats=stats, max_rows=tmax)
    assert fe.launches - n0 == 3             # descriptors (metadata in the kernel parameters), fbank, in-place completion
    torch.cuda.synchronize()
    y = y.cpu().numpy()
    assert y.shape == (len(waves), tmax, 80) and int(stats[160].item()) == sum(frames)
    raws = []
    for i, w in enumerate(waves):
        assert np.all(y[i, frames[i]:] == 0)
        if frames[i] > 1:                   # one frame: std = 0 -> 0/0, no value parity
            ref = F.fbank(w.astype(np.float32), window=tables[0], mel=tables[1])
            raws.append(ref)
            assert np.abs(y[i, :frames[i]] - A.normalization(ref)).max() <= 2e-3
        elif frames[i] == 1:
            raws.append(F.fbank(w.astype(np.float32), window=tables[0], mel=tables[1]))
    allr = np.concatenate(raws).astype(np.float64)
    got = stats.cpu().numpy()
    assert np.allclose(got[:80], allr.sum(0), rtol=1e-4) and np.allclose(got[80:160], (allr ** 2).sum(0), rtol=1e-4)


def test_wav_dither_statistics(fe, tables):
    """kaldi.fbank(dither=d) adds d * N(0,1) to every frame element before DC removal (kaldi.py:179-181).  Stochastic
    (torch's global generator there, Philox here), so the distribution is checked: an all-zero waveform with dither d
    must have the per-bin mean log-mel of a sigma = d white-noise waveform (oracle fbank, no dither); same seed ->
    bit-identical, other seed -> different; frames of one utterance draw independent noise; int16 and fp32 input."""
    d = 3.0
    n = 400 + 160 * 2999                                           # 3000 frames
    zeros = np.zeros(n, np.int16)
    y1, fr = run_raw(fe, [zeros], layout='ragged', wav_dither=d, dither_seed=1234)
    y2, _ = run_raw(fe, [zeros], layout='ragged', wav_dither=d, dither_seed=1234)
    y3, _ = run_raw(fe, [zeros], layout='ragged', wav_dither=d, dither_seed=1235)
    y4, _ = run_raw(fe, [zeros.astype(np.float32)], dtype=np.float32, layout='ragged', wav_dither=d, dither_seed=1234)
    assert fr.tolist() == [3000] and np.isfinite(y1).all()
    assert np.array_equal(y1, y2) and not np.array_equal(y1, y3)
    assert np.array_equal(y1, y4)                                   # same key, same frame elements: same noise
    assert np.abs(y1[0] - y1[1]).max() > 0.1                        # frames are not copies of each other
    noise = np.random.default_rng(0).normal(0.0, d, n).astype(np.float32)
    ref = F.fbank(noise, window=tables[0], mel=tables[1])
    # per-bin std of log-mel over frames is <= 1.3 (1-2 fft bins per mel bin at the low end): 3000 frames -> 0.024
    assert np.abs(y1.mean(0) - ref.mean(0)).max() < 0.15
    assert np.abs(y1.std(0) - ref.std(0)).max() < 0.15
    # a real signal well above the dither level is barely moved
    x = signals.make('speech', 16000, 5)
    a, _ = run_raw(fe, [x], layout='ragged')
    b, _ = run_raw(fe, [x], layout='ragged', wav_dither=1.0, dither_seed=7)
    assert 0 < np.abs(a - b).max() and np.median(np.abs(a - b)) < 0.05
    # fused speed perturb + dither is refused loudly (the dataset mirror resamples first in that case)
    from openeat_b200._lib import FrontendError
    with pytest.raises(FrontendError):
        run_raw(fe, [x], layout='ragged', wav_dither=1.0, speed_ratios=np.array([[9, 10]]))


def test_kernel_timing_hook(fe):
    """oe_frontend_set_kernel_timing / oe_frontend_fbank_kernel_ms: the events the library records around its fbank
    kernel (bench.py's roofline.in_step); refused loudly when no timed call exists."""
    from openeat_b200._lib import FrontendError
    x = [signals.make('white', 160000, 3)] * 8
    fe.set_kernel_timing(True)
    with pytest.raises(FrontendError):
        fe.fbank_kernel_ms()
    y1, _ = run_raw(fe, x, layout='ragged')
    ms = fe.fbank_kernel_ms()
    assert 0.0 < ms < 5.0
    fe.set_kernel_timing(False)
    y2, _ = run_raw(fe, x, layout='ragged')
    assert np.array_equal(y1, y2)
    with pytest.raises(FrontendError):
        fe.fbank_kernel_ms()


def test_custom_mel_weights_use_the_parameter_kernel(tables):
    """A handle whose mel matrix has the standard sparsity structure but other VALUES must not get the kernel that
    carries torchaudio's weights as immediates: it runs the first-generation kernel with the weights as parameters.
    Parity against the oracle with that very matrix; the default handle (mel=None -> the library's own table) is
    torchaudio's matrix bit for bit."""
    from openeat_b200.frontend import Frontend
    win, mel = tables
    scale = (1.0 + 0.01 * np.cos(np.arange(80)))[:, None].astype(np.float32)
    mel2 = (mel * scale).astype(np.float32)                      # same structure, every weight changed
    fe2 = Frontend(mel_bins=80, sample_rate=16000, mel=mel2)
    x = signals.make('speech', 48000, 9)
    y, fr = run_raw(fe2, [x], layout='ragged')
    ref = F.fbank(x.astype(np.float32), window=win, mel=mel2)
    assert np.abs(y - ref).max() <= 1e-3
    base = Frontend(mel_bins=80, sample_rate=16000, torch_tables=False)      # the C library's own tables
    assert np.array_equal(base.tables()[1], mel)
    y0, _ = run_raw(base, [x], layout='ragged')
    assert np.abs(y0 - F.fbank(x.astype(np.float32), window=win, mel=mel)).max() <= 1e-3
    assert np.abs(y - y0).max() > 1e-3                           # the custom weights really were used


@pytest.mark.parametrize('case', ['fbank_mel23_speech_8000', 'fbank_mel23_white_560', 'fbank_mel40_white_8000',
                                  'fbank_mel40_dcsine_8000'])
def test_other_mel_bin_counts_against_torchaudio(golden_dir, case):
    """feature_extraction_conf['mel_bins'] other than 80 (dataset.py:95) runs the table-driven first-generation kernel
    (oe_fbank_kernel<.., false, ..>): torchaudio-generated goldens (oracle/make_golden_r02.py), int16 and fp32 input,
    ragged and padded layout, plus the two-phase chain (per-utterance normalisation, masks) against the oracle."""
    from openeat_b200.frontend import Frontend
    g = np.load(os.path.join(golden_dir, case + '.npz'))
    bins = int(g['mel'].shape[0])
    fe = Frontend(mel_bins=bins, sample_rate=16000)
    assert np.array_equal(fe.tables()[1], g['mel'])                    # torch-built table == torchaudio's, bit for bit
    gap = float(np.abs(g['y32'] - g['y64']).max())
    for dtype in (np.int16, np.float32):
        y, fr = run_raw(fe, [g['pcm']], dtype=dtype, layout='ragged')
        assert fr.tolist() == [g['y32'].shape[0]] and y.shape == g['y32'].shape
        if 'dcsine' in case:
            assert np.abs(y - g['y64']).max() <= max(1e-3, gap)
        else:
            assert np.abs(y - g['y32']).max() <= 1e-3
    # batch of three copies with different lengths, padded, normalised, masked
    waves = [g['pcm'], g['pcm'][:len(g['pcm']) // 2 + 400], g['pcm'][100:]]
    tm = np.array([[[2, 5]], [[0, 1]], [[1, 30]]], np.int32)
    fm = np.array([[[3, 9]], [[0, 2]], [[bins - 4, bins]]], np.int32)
    y, fr = run_raw(fe, waves, layout='padded', normalization=True, tmask=tm, fmask=fm)
    for i, w in enumerate(waves):
        ref = A.normalization(F.fbank(w.astype(np.float32), num_mel_bins=bins, window=fe.tables()[0], mel=g['mel']))
        t = ref.shape[0]
        assert fr[i] == t
        ref = A.apply_spec_augmentation(ref, [(int(tm[i, 0, 0]), min(int(tm[i, 0, 1]), t))], [tuple(fm[i, 0])])
        assert np.array_equal(y[i, :t] == 0, ref == 0)
        if 'dcsine' not in case and t > 2:
            assert np.abs(y[i, :t] - ref).max() <= 5e-3
        assert np.all(y[i, t:] == 0)


def test_processor_chain_with_23_bins(golden_dir):
    """The wenet-style processor chain at compute_fbank's own default (23 bins: 92-byte rows, so per-sample row slices
    are not 16-byte aligned): utt_normalize -> spec_sub -> spec_aug -> global_cmvn, every stage one launch sequence per
    group, against the oracle with the same random draws."""
    from openeat_b200 import processor as P
    from openeat_b200.frontend import Frontend
    from oracle import cmvn as C
    g = np.load(os.path.join(golden_dir, 'fbank_mel23_speech_8000.npz'))
    fe = Frontend(mel_bins=23, sample_rate=16000)
    win = fe.tables()[0]
    pcm = g['pcm']
    samples = [{'key': 'a', 'wav': pcm, 'sample_rate': 16000}, {'key': 'b', 'wav': pcm[:5000], 'sample_rate': 16000},
               {'key': 'c', 'wav': pcm[33:7000], 'sample_rate': 16000}]
    mean, istd = torch.linspace(5.0, 9.0, 23), torch.linspace(0.3, 0.5, 23)
    random.seed(5)
    out = list(P.global_cmvn(P.spec_aug(P.spec_sub(P.utt_normalize(P.compute_fbank(iter(samples))), max_t=10, num_t_sub=2),
                                        num_t_mask=1, num_f_mask=1, max_t=8, max_f=4), mean, istd))
    random.seed(5)
    xs = [A.normalization(F.fbank(s['wav'].astype(np.float32), num_mel_bins=23, window=win, mel=g['mel'])) for s in samples]
    xs = [A.spec_substitute(x, max_t=10, num_t_sub=2) for x in xs]
    xs = [A.spec_augmentation(x, 1, 1, 8, 4) for x in xs]
    for o, x in zip(out, xs):
        ref = C.global_cmvn(x, mean.numpy(), istd.numpy())
        got = o['feat'].cpu().numpy()
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 3e-3


def test_output_canaries_stay_intact(fe):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES are checked the blunt way: the output
    tensor sits between two guard bands of a sentinel bit pattern, for every launch sequence (single pass with padding
    fill, in-place completion, spec_sub through the scratch, ragged rows, odd row counts)."""
    from openeat_b200.frontend import pack_waveforms
    frames = [65, 1, 33, 0, 32, 97]
    lens = [400 + 160 * (m - 1) if m else 300 for m in frames]
    waves = [signals.make('white', n, 300 + i) for i, n in enumerate(lens)]
    buf, offs, ln = pack_waveforms(waves)
    dev = buf.cuda()
    B, G, tmax = len(waves), 257, 101
    sentinel = torch.tensor([0x7FC0DEAD], dtype=torch.int32).view(torch.float32).item()
    mean = torch.linspace(8.0, 12.0, 80, device='cuda')
    istd = torch.linspace(0.4, 0.6, 80, device='cuda')
    maps = [np.arange(m, dtype=np.int32)[::-1].copy() for m in frames]
    modes = [dict(), dict(cmvn=(mean, istd), cmvn_on_padding=True), dict(normalization=True),
             dict(normalization=True, cmvn=(mean, istd), cmvn_on_padding=True, tmask=np.array([[[0, 3]]] * B, np.int32)),
             dict(normalization=True, frame_maps=maps), dict(feature_dither=0.2, dither_seed=3)]
    for kw in modes:
        for layout in ('padded', 'ragged'):
            rows = B * tmax if layout == 'padded' else sum(frames)
            big = torch.empty((rows + 2 * G, 80), device='cuda')
            big.view(torch.int32).fill_(0x7FC0DEAD)
            out = big[G:G + rows]
            if layout == 'padded':
                o_rows, o_n = np.arange(B, dtype=np.int64) * tmax, np.full(B, tmax, np.int32)
            else:
                o_rows, o_n = np.concatenate([[0], np.cumsum(frames[:-1])]).astype(np.int64), np.array(frames, np.int32)
            _, fr = fe.fbank(dev, offs, ln, layout='custom', out=out, out_rows=o_rows, out_nrows=o_n, **kw)
            torch.cuda.synchronize()
            assert fr.tolist() == frames
            guard = torch.cat([big[:G], big[G + rows:]]).view(torch.int32)
            assert bool((guard == 0x7FC0DEAD).all()), (kw.keys(), layout)
            assert not bool((out.view(torch.int32) == 0x7FC0DEAD).any()), (kw.keys(), layout)     # every row was written
    assert sentinel != sentinel                                        # (a NaN pattern: never a legitimate output)
