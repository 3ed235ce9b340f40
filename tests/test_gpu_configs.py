"""BASELINE.json configs at (or near) their full shapes, checked through size-independent properties
(BASELINE.md section 4): streaming windows == whole stream, statistics are linear over shards, spec_sub is a
pure row gather, masks / padding are exact, results are bitwise reproducible."""
import random

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import fbank as F          # noqa: E402
from oracle import signals             # noqa: E402


@pytest.fixture(scope='module')
def fe():
    from openeat_b200.frontend import Frontend
    return Frontend(mel_bins=80, sample_rate=16000)


def device_noise(total, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.randn(total, generator=g, device='cuda') * 3000.0).round_().clamp_(-32768, 32767).to(torch.int16)


def test_config5_twenty_minute_stream_in_16_frame_windows(fe):
    """20-minute stream = 119 998 frames = 7 500 windows of 16 frames (2 800 samples, 240 overlap) run as ONE
    ragged batch, with global CMVN: bitwise equal to the whole-stream features (frames are independent)."""
    n = 160 * 16 * 7500 + 240
    x = device_noise(n, 1006)
    mean = torch.linspace(8.0, 12.0, 80, device='cuda')
    istd = torch.linspace(0.4, 0.6, 80, device='cuda')
    whole, fr = fe.fbank(x, np.array([0]), np.array([n]), layout='ragged', cmvn=(mean, istd))
    assert fr[0] == 119998 + 2                    # 7 500 * 16 frames (+2: the stream length is a multiple of the hop here)
    offs = np.arange(7500, dtype=np.int64) * 2560
    lens = np.full(7500, 2800, dtype=np.int32)
    parts, frames = fe.fbank(x, offs, lens, layout='ragged', cmvn=(mean, istd))
    assert (frames == 16).all()
    assert torch.equal(parts, whole)
    # CPU oracle on a slice in the middle of the stream
    lo = 160 * 16 * 3000
    ref = F.fbank(x[lo:lo + 2800].cpu().numpy().astype(np.float32))
    ref = (ref - mean.cpu().numpy()) * istd.cpu().numpy()
    assert np.abs(parts[16 * 3000:16 * 3001].cpu().numpy() - ref).max() < 1e-3


def test_window_coalescing_is_invisible(monkeypatch):
    """Consecutive windows that continue each other in the waveform buffer and in the output are run as one long
    utterance (full 32-frame tiles instead of half-filled ones): bitwise the same rows and per-window frame counts as with
    OE_NO_COALESCE=1 (statistics to 1e-7); runs are only merged where they really are continuations, and never when something
    per-utterance (normalisation, masks, padded rows) is asked for."""
    from openeat_b200.frontend import Frontend
    monkeypatch.setenv('OE_NO_COALESCE', '1')
    fe_sep = Frontend(mel_bins=80, sample_rate=16000)
    monkeypatch.delenv('OE_NO_COALESCE')
    fe_co = Frontend(mel_bins=80, sample_rate=16000)
    x = device_noise(2560 * 300 + 20000, 77)
    # three streams of 100 / 150 / 49 windows, a stray window in between that is NOT a continuation, one too-short window
    offs, lens = [], []
    for base, count in ((0, 100), (2560 * 100 + 8, 150), (2560 * 251 + 4000, 49)):
        offs += [base + 2560 * i for i in range(count)]
        lens += [2800] * count
    offs.insert(100, 64)
    lens.insert(100, 4000)
    offs.append(8)
    lens.append(300)
    offs, lens = np.array(offs, dtype=np.int64), np.array(lens, dtype=np.int32)
    mean = torch.linspace(8.0, 12.0, 80, device='cuda')
    istd = torch.linspace(0.4, 0.6, 80, device='cuda')
    outs = []
    for f in (fe_sep, fe_co):
        st = torch.zeros(161, dtype=torch.float64, device='cuda')
        n0 = f.launches
        y, fr = f.fbank(x, offs, lens, layout='ragged', cmvn=(mean, istd), stats=st)
        outs.append((y, fr, st, f.launches - n0))
    assert np.array_equal(outs[0][1], outs[1][1]) and outs[0][1][100] == 23 and outs[0][1][-1] == 0
    assert torch.equal(outs[0][0], outs[1][0])
    # the statistics add per-tile fp32 partial sums: other tile boundaries, other roundings (1e-9 relative; the bar is 1e-4)
    assert outs[0][2][160] == outs[1][2][160] and torch.allclose(outs[0][2], outs[1][2], rtol=1e-7, atol=0)
    # per-utterance normalisation: windows must stay separate utterances (each one normalised on its own)
    a, _ = fe_sep.fbank(x, offs[:50], lens[:50], layout='ragged', normalization=True)
    b, _ = fe_co.fbank(x, offs[:50], lens[:50], layout='ragged', normalization=True)
    assert torch.equal(a, b)
    one, _ = fe_co.fbank(x, np.array([0]), np.array([2560 * 49 + 2800]), layout='ragged', normalization=True)
    assert not torch.equal(b, one)


def test_config3_librispeech_shape_sharded_stats_are_linear(fe):
    """Variable 1-35 s utterances, length-sorted dynamic batches (dataset.py:337-352), sharded over 8 'ranks':
    per-rank CMVN statistics add up (bitwise in the count, 1e-12 relative in the sums) to the single-rank pass,
    whatever the batching -- the property the one all-reduce of the path relies on."""
    from openeat_b200.frontend import aligned_offsets
    from openeat_b200.sharding import dynamic_batches, shard_by_length
    lens = signals.lengths_uniform(400, 1.0, 35.0, 1004).astype(np.int32)        # ~2 h of audio
    offs, total = aligned_offsets(lens)
    x = device_noise(total, 1004)
    frames = fe.num_frames_array(lens)

    def stats_of(index_lists):
        st = torch.zeros(161, dtype=torch.float64, device='cuda')
        for idx in index_lists:
            idx = np.asarray(idx)
            fe.fbank(x, offs[idx], lens[idx], layout='ragged', stats=st, want_out=False)
        return st.cpu().numpy()

    everything = stats_of([np.arange(400)])
    batched = stats_of(dynamic_batches(frames.tolist(), 10000))
    shards = shard_by_length(lens, 8)
    per_rank = [stats_of([[int(s[i]) for i in b] for b in dynamic_batches(frames[s].tolist(), 10000)]) for s in shards]
    summed = np.sum(per_rank, axis=0)
    assert everything[160] == batched[160] == summed[160] == frames.sum()
    np.testing.assert_allclose(batched[:160], everything[:160], rtol=1e-11)
    np.testing.assert_allclose(summed[:160], everything[:160], rtol=1e-11)
    loads = [int(lens[s].sum()) for s in shards]
    assert max(loads) - min(loads) <= int(lens.max())
    # oracle VALUES on four utterances (shortest, longest, two in between) of this very list: ragged features, and their
    # sums against a statistics pass over just those four
    from oracle import cmvn as C
    order = np.argsort(lens)
    pick = np.array([order[0], order[133], order[266], order[-1]])
    got, fr4 = fe.fbank(x, offs[pick], lens[pick], layout='ragged')
    got = got.cpu().numpy()
    refs = [F.fbank(x[offs[i]:offs[i] + lens[i]].cpu().numpy().astype(np.float32)) for i in pick]
    assert fr4.tolist() == [r.shape[0] for r in refs]
    # 581 k cells of white noise: the lowest mel bins sit ~40 dB below the frame's energy after pre-emphasis, where two
    # fp32 evaluations (this kernel's FFT, the oracle's pocketfft + matmul) differ by up to a few 1e-3 in the log -- the
    # 1e-3 bound holds for all but a 1e-4 fraction of the cells (and on every golden file, test_gpu_fbank.py)
    d = np.abs(got - np.concatenate(refs))
    assert d.max() <= 4e-3 and (d > 1e-3).mean() <= 1e-4
    s4, q4, n4 = C.compute_cmvn_stats(refs)
    st4 = stats_of([pick])
    assert st4[160] == n4
    np.testing.assert_allclose(st4[:80], s4, rtol=1e-4)
    np.testing.assert_allclose(st4[80:160], q4, rtol=1e-4)
    # mean / variance are sane numbers for Gaussian noise through the mel bank
    mean = everything[:80] / everything[160]
    var = everything[80:160] / everything[160] - mean ** 2
    assert np.all(var > 0) and np.all(np.isfinite(mean))


def test_config4_short_utterances_with_spec_sub_is_a_row_gather(fe):
    """ASRU shape: 256 utterances of 0.5-3 s, spec_sub (3, 30): the substituted batch is exactly the raw batch
    gathered through the composed index map (bitwise), and spec_aug zeros exactly the planned cells."""
    from openeat_b200 import planner
    from openeat_b200.frontend import aligned_offsets
    lens = signals.lengths_uniform(256, 0.5, 3.0, 1005).astype(np.int32)
    offs, total = aligned_offsets(lens)
    x = device_noise(total, 1005)
    frames = fe.num_frames_array(lens)
    raw, _ = fe.fbank(x, offs, lens, layout='padded')
    random.seed(1005)
    fmap, tm, fm = planner.plan_augment(frames, 80, dict(num_t_sub=3, max_t=30),
                                        dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10))
    starts = np.concatenate([[0], np.cumsum(frames[:-1].astype(np.int64))])
    maps = [fmap[s:s + t] for s, t in zip(starts, frames)]
    sub, _ = fe.fbank(x, offs, lens, layout='padded', frame_maps=maps)
    both, _ = fe.fbank(x, offs, lens, layout='padded', frame_maps=maps, tmask=tm, fmask=fm)
    raw, sub, both = raw.cpu().numpy(), sub.cpu().numpy(), both.cpu().numpy()
    from oracle import augment as A
    for b in (0, 1, 17, 100, 255):
        t = int(frames[b])
        # oracle VALUES: raw log-mel of this utterance, then the reference's own substitution / masking arithmetic
        ref = F.fbank(x[offs[b]:offs[b] + lens[b]].cpu().numpy().astype(np.float32))
        assert ref.shape[0] == t and np.abs(raw[b, :t] - ref).max() <= 1e-3
        ref_both = A.apply_spec_augmentation(ref[maps[b]], [tuple(r) for r in tm[b]], [tuple(r) for r in fm[b]])
        assert np.array_equal(both[b, :t] == 0, ref_both == 0) and np.abs(both[b, :t] - ref_both).max() <= 1e-3
        assert np.array_equal(sub[b, :t], raw[b, :t][maps[b]])
        expect = sub[b, :t].copy()
        for s, e in tm[b]:
            expect[s:e] = 0
        for s, e in fm[b]:
            expect[:, s:e] = 0
        assert np.array_equal(both[b, :t], expect)
        assert np.all(both[b, t:] == 0)


def test_config2_full_batch_properties(fe):
    """Batch 256 of 2-10 s with fused speed perturb + normalisation + spec_aug + CMVN: bitwise reproducible,
    padding exactly (0 - mean) * istd, masked cells likewise, every utterance normalised (mean 0, var 1)."""
    from openeat_b200 import planner
    from openeat_b200.frontend import aligned_offsets
    lens = signals.lengths_uniform(256, 2.0, 10.0, 1002).astype(np.int32)
    offs, total = aligned_offsets(lens)
    x = device_noise(total, 1002)
    rs = np.random.default_rng(1003).integers(0, 3, 256)
    ratios = np.array([[(9, 10), (0, 0), (11, 10)][i] for i in rs])
    eff = np.where(ratios[:, 0] > 0, -(-lens.astype(np.int64) * ratios[:, 1] // np.maximum(ratios[:, 0], 1)), lens)
    frames = fe.num_frames_array(eff)
    random.seed(7)
    _, tm, fm = planner.plan_augment(frames, 80, None, dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10))
    mean = torch.linspace(8.0, 12.0, 80, device='cuda')
    istd = torch.linspace(0.4, 0.6, 80, device='cuda')
    kw = dict(layout='padded', normalization=True, tmask=tm, fmask=fm, cmvn=(mean, istd), cmvn_on_padding=True,
              speed_ratios=ratios)
    a, fr = fe.fbank(x, offs, lens, **kw)
    b, _ = fe.fbank(x, offs, lens, **kw)
    assert fr.tolist() == frames.tolist()
    assert torch.equal(a, b)
    plain, _ = fe.fbank(x, offs, lens, layout='padded', normalization=True, speed_ratios=ratios)
    a, plain = a.cpu().numpy(), plain.cpu().numpy()
    zero_after = ((np.float32(0) - mean.cpu().numpy()) * istd.cpu().numpy()).astype(np.float32)
    for i in (0, 3, 128, 255):
        t = int(frames[i])
        assert np.array_equal(a[i, t:], np.broadcast_to(zero_after, a[i, t:].shape))
        for s, e in tm[i]:
            assert np.array_equal(a[i, s:e], np.broadcast_to(zero_after, a[i, s:e].shape))
        for s, e in fm[i]:
            assert np.array_equal(a[i, :t, s:e], np.broadcast_to(zero_after[s:e], a[i, :t, s:e].shape))
        np.testing.assert_allclose(plain[i, :t].mean(0), 0.0, atol=2e-4)
        np.testing.assert_allclose(plain[i, :t].std(0), 1.0, atol=2e-4)
        # oracle VALUES of the whole chain for this utterance: substitute speed oracle -> fbank -> _normalization ->
        # masks -> GlobalCMVN (the resampler tolerance dominates: 0.05 on the int16 scale; the ill-conditioned top bins
        # of speed 0.9 are bounded separately in test_gpu_fbank.py)
        from oracle import augment as A, cmvn as C, speed as S
        w = x[offs[i]:offs[i] + lens[i]].cpu().numpy().astype(np.float32)
        if ratios[i, 0]:
            w = S.resample(w, int(ratios[i, 0]), int(ratios[i, 1]))
        ref = A.normalization(F.fbank(np.asarray(w, np.float32)))
        assert ref.shape[0] == t
        hi = 70 if ratios[i, 0] == 9 else 80
        assert np.abs(plain[i, :t, :hi] - ref[:, :hi]).max() <= 5e-3
        ref = C.global_cmvn(A.apply_spec_augmentation(ref, [tuple(r) for r in tm[i]], [tuple(r) for r in fm[i]]),
                            mean.cpu().numpy(), istd.cpu().numpy())
        assert np.abs(a[i, :t, :hi] - ref[:, :hi]).max() <= 5e-3


def test_multi_tile_ctas_equal_one_utterance_at_a_time():
    """~1000 tiles on 296 persistent CTAs: every CTA walks several tiles, so every cross-tile shared-memory reuse of
    the fbank kernels (raw buffer, exchange areas / power slices, output tile, descriptors, masks) is exercised.
    Frames are independent of their batch, so the batched result must be BITWISE equal to running each utterance
    alone (one tile per CTA, nothing reused) -- plain int16, fp32 input, and the fused speed-perturb + per-utterance
    normalisation + masks + CMVN chain.  (compute-sanitizer's racecheck is not available on the GPU pool; this is
    the race detector.)"""
    from openeat_b200.frontend import Frontend, pack_waveforms
    fe = Frontend(mel_bins=80, sample_rate=16000)
    rng = np.random.default_rng(7)
    B = 40
    lens = rng.integers(5 * 16000, 10 * 16000, B)
    waves = [rng.integers(-3000, 3000, n).astype(np.int16) for n in lens]
    buf, offs, ln = pack_waveforms(waves)
    dev = buf.cuda()
    ratios = np.array([[(0, 0), (9, 10), (11, 10)][i % 3] for i in range(B)])
    mean = torch.linspace(8.0, 12.0, 80, device='cuda')
    istd = torch.linspace(0.4, 0.6, 80, device='cuda')
    tm = np.array([[[3, 9], [40, 70]]] * B, np.int32)
    fm = np.array([[[10, 14]]] * B, np.int32)
    modes = [
        (dev, dict()),
        (dev.float(), dict()),
        (dev, dict(tmask=tm, fmask=fm, cmvn=(mean, istd), cmvn_on_padding=False)),
        (dev, dict(normalization=True, speed_ratios=ratios, tmask=tm, fmask=fm, cmvn=(mean, istd), cmvn_on_padding=False)),
    ]
    for rep in range(2):                                    # twice: a race need not show on the first run
        for wav, kw in modes:
            full, frames = fe.fbank(wav, offs, ln, layout='padded', **kw)
            assert int(np.ceil(frames / 32).sum()) > 3 * 296
            full = full.cpu().numpy()
            for i in range(0, B, 3):                        # every third utterance alone
                kw1 = {k: (v[i:i + 1] if k in ('tmask', 'fmask', 'speed_ratios') else v) for k, v in kw.items()}
                one, f1 = fe.fbank(wav, offs[i:i + 1], ln[i:i + 1], layout='padded', **kw1)
                assert f1[0] == frames[i]
                assert np.array_equal(full[i, :frames[i]], one.cpu().numpy()[0]), (rep, i, sorted(kw))
