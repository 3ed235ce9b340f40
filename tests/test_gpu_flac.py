"""GPU decoder for FLAC lists (csrc/oe_flac_gpu.cuh) against the host decoder / the PCM that was encoded: bit-exact, and
end to end through the reference-facing collate (dataset.py:55-118 for .flac entries)."""
import random

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from test_flac import flac_stream_matrix, lib_encode, speechlike   # noqa: E402
from oracle import flac as oflac                                    # noqa: E402

CONF = {'resample_rate': 16000, 'speed_perturb_rate': 0, 'speeds': [0.9, 1.1, 0.1], 'wav_dither': 0.0, 'mel_bins': 80}


def test_kernel_decodes_every_stream_bit_exactly(tmp_path):
    from openeat_b200.ingest import FlacGpuIngest
    rng = np.random.default_rng(22)
    streams = flac_stream_matrix(rng)
    entries, want = [], []
    for name, (data, pcm) in streams.items():
        f = tmp_path / (name + '.flac')
        f.write_bytes(data)
        entries.append(str(f))
        want.append(pcm)
    for name, s, e in (('enc4096', 0.25, 0.75), ('enc4096', 0.3, 5.0), ('enc1152', 0.07201, 0.07207), ('oracle8', 0.01, 0.1)):
        entries.append('%s,%r,%r' % (tmp_path / (name + '.flac'), s, e))
        pcm = streams[name][1]
        a = int(s * 16000)
        want.append(pcm[a:a + max(0, min(int(e * 16000) - a, len(pcm) - a))])
    ing = FlacGpuIngest(threads=3, ring=2)
    for verify in (True, False):
        b = ing.pack(entries)
        assert b.loaded.all() and b.lens.tolist() == [len(w) for w in want]
        pcm = b.to_device(torch.device('cuda:0'), verify_crc=verify).cpu().numpy()
        b.check()
        for i, w in enumerate(want):
            assert np.array_equal(pcm[b.offsets[i]:b.offsets[i] + b.lens[i]], w), entries[i]


def test_kernel_flags_corruption_and_out_of_range_streams(tmp_path):
    from openeat_b200._lib import FrontendError
    from openeat_b200.ingest import FlacGpuIngest
    rng = np.random.default_rng(23)
    x = speechlike(rng, 12000)
    good = lib_encode(x, block=1152)
    _, pos = oflac.parse_streaminfo(good)
    ing = FlacGpuIngest(threads=2, ring=2)
    (tmp_path / 'good.flac').write_bytes(good)
    names, n_bad = [str(tmp_path / 'good.flac')], 0
    for trial in range(24):
        bad = bytearray(good)
        at = int(rng.integers(pos + 8, len(good) - 2))
        bad[at] ^= 1 << int(rng.integers(0, 8))
        p = tmp_path / ('bad%d.flac' % trial)
        p.write_bytes(bytes(bad))
        names.append(str(p))
    y = speechlike(rng, 2000).astype(np.int64)
    (tmp_path / 'o32.flac').write_bytes(oflac.encode(y[None], 16000, 16, kind='lpc', order=32, lpc=(8, 7, rng.integers(-50, 51, 32).tolist())))
    names.append(str(tmp_path / 'o32.flac'))
    b = ing.pack(names, report=False)
    pcm = b.to_device(torch.device('cuda:0')).cpu().numpy()
    err = b.errors.numpy()
    assert err[0] == 0 and np.array_equal(pcm[b.offsets[0]:b.offsets[0] + b.lens[0]], x)
    assert err[len(names) - 1] == 4
    for i in range(1, len(names) - 1):
        assert (not b.loaded[i]) or err[i] != 0                   # refused by the header walk or flagged by the kernel
        n_bad += int(err[i] != 0)
    assert n_bad >= 18
    with pytest.raises(FrontendError, match='CRC-16|does not end'):
        b.check()


def test_collate_of_flac_files_equals_collate_of_their_pcm(tmp_path, capsys):
    """audio_collate_func on a .flac list (GPU decode) == collate_packed on the same PCM, bitwise; a corrupted file is
    printed and dropped like any unreadable file (dataset.py:108-111); the pipelined path gives the same batches."""
    from openeat_b200.dataset import PrefetchingCollator, audio_collate_func
    from openeat_b200.frontend import aligned_offsets
    from openeat_b200.ingest import flac_gpu_batches
    rng = np.random.default_rng(31)
    lens = [16000, 9000, 52000, 12345, 7777, 30001]
    pcm = [speechlike(rng, n) for n in lens]
    batch = []
    for i, x in enumerate(pcm):
        p = tmp_path / ('u%d.flac' % i)
        p.write_bytes(lib_encode(x, block=4096 if i % 2 else 1152))
        batch.append(('utt%d' % i, str(p), [i + 1] * (i + 2), 1.0))
    batch.append(('seg', str(tmp_path / 'u2.flac') + ',0.25,1.75', [9, 9], 1.0))
    pcm.append(pcm[2][4000:28000])
    kw = dict(normalization=True, spec_aug=True, spec_aug_conf=dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10))
    fn = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, **kw)
    random.seed(77)
    keys, out = fn([batch])
    ln = np.array([len(x) for x in pcm], dtype=np.int32)
    offs, total = aligned_offsets(ln)
    buf = np.zeros(total, dtype=np.int16)
    for o, x in zip(offs, pcm):
        buf[o:o + len(x)] = x
    random.seed(77)
    keys2, out2 = fn.collate_packed(torch.from_numpy(buf).cuda(), offs, ln, [b[0] for b in batch], [b[2] for b in batch],
                                    [b[3] for b in batch])
    assert list(keys) == list(keys2)
    assert torch.equal(out['features'], out2['features']) and torch.equal(out['features_length'], out2['features_length'])
    assert torch.equal(out['targets'], out2['targets'])
    # pipelined: three batches through flac_gpu_batches + PrefetchingCollator
    random.seed(77)
    got = list(PrefetchingCollator(fn, flac_gpu_batches([batch, batch[:3], batch])))
    random.seed(77)
    ref = [fn([batch]), fn([batch[:3]]), fn([batch])]
    for (k1, o1), (k2, o2) in zip(got, ref):
        assert list(k1) == list(k2) and torch.equal(o1['features'], o2['features'])
    # a flipped bit inside a frame body: that utterance is dropped, the others are unchanged
    data = bytearray((tmp_path / 'u3.flac').read_bytes())
    data[len(data) // 2] ^= 0x10
    (tmp_path / 'u3.flac').write_bytes(bytes(data))
    capsys.readouterr()
    random.seed(77)
    keys3, out3 = fn([batch])
    assert 'utt3' not in list(keys3) and len(keys3) == len(keys) - 1
    assert 'rejected by the GPU decoder' in capsys.readouterr().out


def test_batch_without_a_decodable_file(tmp_path, capsys):
    """Every entry unreadable (missing file, damaged header): the batch comes back empty through both entry points instead
    of failing (dataset.py:108-111 drops unreadable utterances one by one)."""
    from openeat_b200.dataset import PrefetchingCollator, audio_collate_func
    from openeat_b200.ingest import flac_gpu_batches
    (tmp_path / 'bad.flac').write_bytes(b'fLaC' + bytes(64))
    batch = [('a', str(tmp_path / 'bad.flac'), [1], 1.0), ('b', str(tmp_path / 'missing.flac'), [2], 1.0)]
    fn = audio_collate_func(data_type='wav', feature_extraction_conf=CONF)
    keys, out = fn([batch])
    assert len(keys) == 0
    got = list(PrefetchingCollator(fn, flac_gpu_batches([batch, batch])))
    assert len(got) == 2 and all(len(k) == 0 for k, _ in got)
    assert 'STREAMINFO' in capsys.readouterr().out
