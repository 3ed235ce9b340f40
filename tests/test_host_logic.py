"""CPU tests of the host-side logic: RNG plans, sharding, CMVN file formats, the world_size-2 gloo path."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from oracle import augment as A


def test_plans_follow_the_reference_rng_order():
    from openeat_b200.feature_processor import plan_spec_augmentation, plan_spec_substitute
    for seed, T in [(1, 498), (2, 48), (3, 1), (4, 7)]:
        random.seed(seed)
        subs = A.plan_spec_substitute(T, max_t=30, num_t_sub=3)
        ref_aug = A.plan_spec_augmentation(T, 80, 3, 2, 50, 10)
        tail = random.random()
        random.seed(seed)
        idx = plan_spec_substitute(T, max_t=30, num_t_sub=3)
        aug = plan_spec_augmentation(T, 80, 3, 2, 50, 10)
        assert np.array_equal(idx, A.substitute_index_map(T, subs)) and aug == ref_aug
        assert random.random() == tail                              # same number of RNG words consumed


def test_native_planner_continues_python_random():
    """oe_plan_speeds / oe_plan_augment (C, Mersenne Twister) == the reference's Python loops for the same
    random.seed, and they leave Python's generator in the same state."""
    from openeat_b200 import planner
    from oracle import speed as S
    for seed, rate, cfg in [(0, 0.5, None), (1, 1.0, [0.9, 1.1, 0.1]), (2, 0.7, [0.9, 1.1, 0]), (3, 0.3, [1.05]), (4, 0.0, None)]:
        frames = np.array([498, 48, 7, 1, 298, 1000, 33, 2], np.int32)
        item = np.array([1.0, 0.9, 1.1, 1.0, 1.0, 1.0, 0.9, 1.0])
        active = np.array([1, 1, 0, 1, 1, 1, 1, 1], bool)
        random.seed(seed)
        ref_speed = []
        for i in range(8):
            s = item[i]
            if active[i] and random.random() < rate:
                s = S.speed_generator(cfg)
            ref_speed.append(s)
        subs = [A.plan_spec_substitute(int(t), max_t=30, num_t_sub=3) for t in frames]
        ref_map = np.concatenate([A.substitute_index_map(int(t), s) for t, s in zip(frames, subs)])
        ref_aug = [A.plan_spec_augmentation(int(t), 80, 3, 2, 50, 10) for t in frames]
        tail = random.random()
        random.seed(seed)
        speed = planner.plan_speeds(rate, cfg, item, active)
        fmap, tm, fm = planner.plan_augment(frames, 80, dict(max_t=30, num_t_sub=3),
                                            dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10))
        assert speed.tolist() == ref_speed
        assert np.array_equal(fmap, ref_map)
        assert tm.tolist() == [[list(r) for r in p[0]] for p in ref_aug]
        assert fm.tolist() == [[list(r) for r in p[1]] for p in ref_aug]
        assert random.random() == tail
    # defaults of feature_processor.py when a conf omits keys; aug only; sub only
    random.seed(9)
    ref = A.plan_spec_augmentation(100, 80)
    random.seed(9)
    _, tm, fm = planner.plan_augment([100], 80, None, {})
    assert tm[0].tolist() == [list(r) for r in ref[0]] and fm[0].tolist() == [list(r) for r in ref[1]]
    assert planner.plan_augment([5, 6], 80, None, None) == (None, None, None)


def test_speed_generator_is_pure_host_logic(golden_dir):
    from openeat_b200.audio_processor import _speed_generator
    a = np.load(os.path.join(golden_dir, 'augment.npz'))
    random.seed(5)
    draws = [_speed_generator([0.9, 1.1, 0.1]) for _ in range(8)] + [_speed_generator(None)] + [_speed_generator([1.05])]
    assert np.array_equal(np.array(draws), a['speed_draws'])
    with pytest.raises(AssertionError):
        _speed_generator([1.1, 0.9, 0.1])


def test_cmvn_file_formats(golden_dir, tmp_path):
    from openeat_b200.cmvn import load_cmvn, write_json_cmvn
    g = np.load(os.path.join(golden_dir, 'cmvn.npz'))
    mean, istd = load_cmvn(os.path.join(golden_dir, 'cmvn_stats.json'), True)
    assert np.array_equal(mean, g['mean_json']) and np.array_equal(istd, g['istd_json'])
    mean, istd = load_cmvn(os.path.join(golden_dir, 'cmvn_stats.kaldi.txt'), False)
    assert np.array_equal(mean, g['mean_kaldi']) and np.array_equal(istd, g['istd_kaldi'])
    p = str(tmp_path / 'c.json')
    write_json_cmvn(p, g['sum'], g['sumsq'], int(g['count']))
    mean2, istd2 = load_cmvn(p, True)
    assert np.array_equal(mean2, g['mean_json']) and np.array_equal(istd2, g['istd_json'])
    (tmp_path / 'bin').write_text('\0B junk')
    with pytest.raises(ValueError):
        load_cmvn(str(tmp_path / 'bin'), False)


def test_sharding_and_batching():
    from openeat_b200.sharding import dynamic_batches, shard_by_length, static_batches
    rng = np.random.default_rng(1004)
    lens = np.round(rng.uniform(1, 35, 2000) * 16000).astype(np.int64)
    for ws in (1, 2, 4, 8):
        shards = shard_by_length(lens, ws)
        assert sorted(np.concatenate(shards).tolist()) == list(range(2000))
        loads = np.array([lens[s].sum() for s in shards])
        assert loads.max() - loads.min() <= lens.max()              # balanced to within one utterance
        assert all(np.array_equal(a, b) for a, b in zip(shards, shard_by_length(lens, ws)))   # deterministic
    frames = (1 + (lens - 400) // 160).tolist()
    batches = dynamic_batches(frames, 10000)
    assert sorted(i for b in batches for i in b) == list(range(2000))
    assert all(sum(frames[i] for i in b) <= 10000 or len(b) == 1 for b in batches)
    assert static_batches(5, 2) == [[0, 1], [2, 3], [4]]


WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from openeat_b200.cmvn import all_reduce_stats, write_json_cmvn, load_cmvn
from openeat_b200.sharding import shard_by_length
from oracle import fbank as F, signals, cmvn as C
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', init_method='env://')
lens = [8000, 4000, 12000, 6000, 9000, 5000, 700]
shard = shard_by_length(lens, world)[rank]
feats = [F.fbank(signals.make('speech', lens[i], 500 + i).astype(np.float32)) for i in shard]
s, q, n = C.compute_cmvn_stats(feats)                 # the local (per-GPU) part, stood in for by the oracle on CPU
stats = torch.from_numpy(np.concatenate([s, q, [float(n)]]))
all_reduce_stats(stats)                               # the path's one collective
if rank == 0:
    st = stats.numpy()
    write_json_cmvn(sys.argv[2], st[:80], st[80:160], int(round(st[160])))
dist.barrier()
dist.destroy_process_group()
'''


def test_world_size_2_stats_allreduce_gloo(tmp_path):
    """N>1 host path on CPU: shard by length, local stats, ONE all-reduce of 161 doubles, rank 0 writes the
    JSON the reference's load_cmvn reads; result equals the single-process statistics."""
    from oracle import cmvn as C
    from oracle import fbank as F
    from oracle import signals
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    out = str(tmp_path / 'cmvn.json')
    port = 29500 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), root, out], env=env))
    for p in procs:
        assert p.wait(timeout=240) == 0
    lens = [8000, 4000, 12000, 6000, 9000, 5000, 700]
    feats = [F.fbank(signals.make('speech', n, 500 + i).astype(np.float32)) for i, n in enumerate(lens)]
    s, q, n = C.compute_cmvn_stats(feats)
    import json
    st = json.load(open(out))
    assert st['frame_num'] == n
    np.testing.assert_allclose(st['mean_stat'], s, rtol=1e-12)
    np.testing.assert_allclose(st['var_stat'], q, rtol=1e-12)
    mean, istd = C.load_cmvn(out, True)
    assert np.isfinite(mean).all() and np.isfinite(istd).all()


def test_audio_dataset_matches_reference_batches(golden_dir):
    """List parsing + length filter + offline speed list + sort + static / dynamic / shuffle batching against the
    batches the reference's own AudioDataset built from the same format.data (tests/golden/audio_dataset.json)."""
    import json
    from openeat_b200.dataset import AudioDataset
    gold = json.load(open(os.path.join(golden_dir, 'audio_dataset.json'), encoding='utf-8'))
    char_dict = json.load(open(os.path.join(golden_dir, 'format.dict.json'), encoding='utf-8'))
    path = os.path.join(golden_dir, 'format.data')
    cases = [('static', dict(batch_type='static', batch_size=4, sort=True, max_length=2000, min_length=10)),
             ('dynamic', dict(batch_type='dynamic', max_frames_in_batch=3000, sort=True, max_length=2000, min_length=10)),
             ('dynamic_unsorted_speed', dict(batch_type='dynamic', max_frames_in_batch=1500, sort=False,
                                             speed_perturb=True, max_length=2000, min_length=10)),
             ('shuffle', dict(batch_type='shuffle', batch_size=8, sort=False))]
    for tag, kw in cases:
        ds = AudioDataset(path, char_dict, None, data_type='wav', **kw)
        got = [[[x[0], x[1], list(x[2]), float(x[3])] for x in (b if tag != 'shuffle' else [b])] for b in ds.data]
        assert got == gold[tag], tag
        assert ds.batch_size == gold[tag + '_batch_size'] and len(ds) == len(gold[tag])
    ds = AudioDataset(path, char_dict, None, data_type='wav', batch_type='static', batch_size=4)
    assert any(',1.5,4.25' in item[1] for b in ds.data for item in b)        # segmented entries survive as path,start,end
    with pytest.raises(IndexError):                # a wav list read as a Kaldi-feature list: feat_shape has no ",dim" -- the reference raises the same
        AudioDataset(path, char_dict, None, data_type='kaldi')


def test_kaldi_io_roundtrip_and_rxspecifier(tmp_path):
    """openeat_b200.kaldi_io: the binary matrix format kaldi_io.read_mat (dataset.py:138) reads -- archive written
    here, read back by 'path:offset' (feats.scp form), by plain path (first entry) and as a generator; a double
    matrix and a malformed header."""
    import struct
    from openeat_b200 import kaldi_io
    rng = np.random.default_rng(0)
    mats = {'utt%d' % i: rng.normal(size=(3 + 7 * i, 80)).astype(np.float32) for i in range(4)}
    ark = str(tmp_path / 'feats.ark')
    scp = kaldi_io.write_mat_ark(ark, mats.items())
    for k, m in mats.items():
        assert np.array_equal(kaldi_io.read_mat(scp[k]), m)
    assert np.array_equal(kaldi_io.read_mat(ark), mats['utt0'])
    assert [(k, v.shape) for k, v in kaldi_io.read_mat_ark(ark)] == [(k, m.shape) for k, m in mats.items()]
    dm = str(tmp_path / 'd.ark')
    with open(dm, 'wb') as fd:
        fd.write(b'x \0BDM \4' + struct.pack('<i', 2) + b'\4' + struct.pack('<i', 3) + np.arange(6, dtype='<f8').tobytes())
    assert np.array_equal(kaldi_io.read_mat(dm), np.arange(6.0).reshape(2, 3))
    bad = str(tmp_path / 'bad.ark')
    with open(bad, 'wb') as fd:
        fd.write(b'x \0BCM ' + b'\0' * 32)
    with pytest.raises(ValueError):
        kaldi_io.read_mat(bad)


def test_tar_shard_reader(tmp_path):
    """processor.tar_file_and_group: wenet shard tars (<key>.wav + <key>.txt members) -> sample dicts; an entry whose
    wav cannot be decoded is skipped, the rest of the shard survives."""
    import io
    import tarfile
    import wave
    from openeat_b200 import processor
    rng = np.random.default_rng(1)
    waves = {'a%d' % i: rng.integers(-3000, 3000, 1600 + 100 * i).astype(np.int16) for i in range(3)}
    shard = str(tmp_path / 'shard_000.tar')
    with tarfile.open(shard, 'w') as tar:
        def add(name, blob):
            ti = tarfile.TarInfo(name)
            ti.size = len(blob)
            tar.addfile(ti, io.BytesIO(blob))
        for k, w in waves.items():
            buf = io.BytesIO()
            with wave.open(buf, 'wb') as f:
                f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(w.tobytes())
            add(k + '.txt', ('text of ' + k).encode())
            add(k + '.wav', buf.getvalue())
        add('broken.txt', b'x')
        add('broken.wav', b'not a wav file')
    got = list(processor.tar_file_and_group([shard]))
    assert [s['key'] for s in got] == list(waves)
    for s in got:
        assert np.array_equal(s['wav'], waves[s['key']]) and s['sample_rate'] == 16000 and s['txt'] == 'text of ' + s['key']


def test_native_ingest_matches_the_stdlib_reader(tmp_path):
    """oe_ingest_probe / oe_ingest_read (dataset.py:55-75 natively): mono and stereo (channel 0) 16-bit PCM, segmented
    entries 'path,start,end' incl. one that runs past the end of the file, 8-sample aligned packing -- equal to the
    stdlib-`wave` reader sample for sample; what this ingest cannot fill an int16 buffer from (32-bit samples, a damaged
    FLAC stream; sound FLAC files: tests/test_flac.py) and a missing file are reported with an explicit reason and marked not loaded (never silently dropped)."""
    import wave
    from openeat_b200.dataset import read_wav
    from openeat_b200.ingest import NativeIngest
    rng = np.random.default_rng(0)

    def wr(path, pcm, ch=1, width=2, sr=16000):
        with wave.open(str(path), 'wb') as w:
            w.setnchannels(ch)
            w.setsampwidth(width)
            w.setframerate(sr)
            w.writeframes(pcm.tobytes())
    a = rng.integers(-3000, 3000, 16001).astype('<i2')
    b = rng.integers(-3000, 3000, (9000, 2)).astype('<i2')
    wr(tmp_path / 'a.wav', a)
    wr(tmp_path / 'b.wav', b, ch=2, sr=8000)
    wr(tmp_path / 'c.wav', rng.integers(-3000, 3000, 500).astype('<i4'), width=4)
    (tmp_path / 'd.flac').write_bytes(b'fLaC' + b'\0' * 64)
    p = str(tmp_path)
    entries = [p + '/a.wav', p + '/b.wav', p + '/a.wav,0.25,0.75', p + '/c.wav', p + '/d.flac', p + '/missing.wav',
               p + '/a.wav,0.9,5.0']
    ing = NativeIngest(threads=3, ring=2)
    for _ in range(3):                                            # walks the ring
        buf, offs, lens, rates, loaded, slot = ing.load(entries, keys=['k%d' % i for i in range(7)])
        x = buf.numpy()
        assert loaded.tolist() == [True, True, True, False, False, False, True]
        assert rates.tolist()[:3] == [16000, 8000, 16000] and (offs % 8 == 0).all()
        assert lens.tolist() == [16001, 9000, 8000, 0, 0, 0, 1601]
        assert np.array_equal(x[offs[0]:offs[0] + lens[0]], a)
        assert np.array_equal(x[offs[1]:offs[1] + lens[1]], b[:, 0])
        for i, (s, e) in ((2, (0.25, 0.75)), (6, (0.9, 5.0))):
            ref, _ = read_wav(p + '/a.wav', s, e)
            assert np.array_equal(x[offs[i]:offs[i] + lens[i]], ref)
    errs = [ing.lib.oe_ingest_error(ing.handle, i).decode() for i in range(7)]
    assert '32-bit' in errs[3] and 'STREAMINFO' in errs[4] and 'No such file' in errs[5] and errs[0] == ''


def test_audio_dataset_parses_kaldi_feature_lists(tmp_path):
    """AudioDataset(data_type='kaldi') (dataset.py:325-331): ``feat:file.ark:offset`` and ``feat_shape:frames,dim``."""
    from openeat_b200.dataset import AudioDataset
    lines = ['utt:u%d\tfeat:/data/f.ark:%d\tfeat_shape:%d,80\ttext:a b' % (i, 17 + 100 * i, t)
             for i, t in enumerate([120, 30, 500, 75])]
    f = tmp_path / 'format.data'
    f.write_text('\n'.join(lines) + '\n')
    ds = AudioDataset(str(f), {'A': 1, 'B': 2, '<unk>': 0, ' ': 3}, data_type='kaldi', batch_type='static', batch_size=2,
                      sort=True, max_length=400, min_length=40)
    assert ds.input_size == 80
    assert [[it[0] for it in b] for b in ds] == [['u3', 'u0']]    # 30 and 500 frames are filtered out, sorted by length
    assert ds[0][0][1] == '/data/f.ark:317'


def test_asynchronous_ingest_batches(tmp_path):
    """ingest_batches: batches read ahead by the handle's native driver thread (oe_ingest_submit / oe_ingest_wait) arrive
    in order and equal the files; a ring slot that is too small is replaced transparently; failures are reported."""
    import wave
    from openeat_b200.ingest import NativeIngest, ingest_batches
    rng = np.random.default_rng(1)
    data, batches = {}, []
    for b in range(5):
        items = []
        for u in range(6):
            x = rng.integers(-3000, 3000, int(rng.integers(500, 60000) * (8 if b == 3 else 1))).astype('<i2')
            p = str(tmp_path / ('b%du%d.wav' % (b, u)))
            with wave.open(p, 'wb') as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(16000)
                w.writeframes(x.tobytes())
            data[p] = x
            items.append(('k%d_%d' % (b, u), p, [1, 2], 1.0))
        items.append(('bad%d' % b, str(tmp_path / 'nope.wav'), [1], 1.0))
        batches.append(items)
    ing = NativeIngest(threads=2, ring=5)
    ing._seen = 1 << 16                                            # small first guess: batch 3 (8x longer) outgrows its slot
    seen = []
    for buf, offs, lens, keys, labels, speeds, rates, loaded, release in ingest_batches(iter(batches), ing, depth=2):
        x = buf.numpy()
        seen.append(keys[0])
        assert loaded.tolist() == [True] * 6 + [False] and lens[6] == 0 and (offs % 8 == 0).all()
        for i in range(6):
            p = batches[len(seen) - 1][i][1]
            assert np.array_equal(x[offs[i]:offs[i] + lens[i]], data[p])
        release(None)
    assert seen == ['k%d_0' % b for b in range(5)]


def test_host_pad_rows(tmp_path):
    """oe_host_pad_rows (pad_sequence of dataset.py:214-218 on the reader pool): ragged rows -> (B, Tmax, F), padding rows
    zero or a given row; odd row sizes and unaligned destinations take the head / tail paths of the streaming copy."""
    torch = pytest.importorskip('torch')
    from openeat_b200.ingest import NativeIngest
    ing = NativeIngest(threads=4)
    rng = np.random.default_rng(5)
    for F, frames in ((80, [300, 129, 128, 1, 0]), (23, [7, 7, 3]), (3, [1000, 999])):
        tmax = max(frames)
        src = torch.from_numpy(rng.normal(size=(sum(frames), F)).astype(np.float32))
        for pad_row in (None, rng.normal(size=F).astype(np.float32)):
            raw = torch.full((len(frames) * tmax * F + 5,), 7.0)
            dst = raw[3:3 + len(frames) * tmax * F].view(len(frames), tmax, F)      # 12-byte offset: unaligned rows
            ing.pad_rows(src, frames, tmax, dst, pad_row)
            want = np.zeros((len(frames), tmax, F), np.float32) if pad_row is None else np.broadcast_to(pad_row, (len(frames), tmax, F)).copy()
            o = 0
            for b, n in enumerate(frames):
                want[b, :n] = src[o:o + n].numpy()
                o += n
            assert np.array_equal(dst.numpy(), want)
            assert raw[:3].eq(7).all() and raw[-2:].eq(7).all()
    with pytest.raises(Exception):
        ing.pad_rows(src, [5000], 10, dst)                                          # frames > tmax
