"""FLAC streams (the LibriSpeech recipe's corpus; the reference reads them through torchaudio.load / libsox,
openeat/dataset/dataset.py:62-72).  libFLAC is third-party code absent from the reference tree and from this image, so
the golden material is the format's own: the worked examples of RFC 9639 appendix D (bytes below; their CRC-8, CRC-16
and MD5 signature all check out, which also proves the bytes are the published ones), plus streams written by the
oracle's configurable encoder so that every branch of the format is met.  CPU only: the decoder is host code inside
libopeneat_frontend.so (csrc/oe_flac.h), reached through the C ABI (oe_flac_info / oe_flac_decode / oe_ingest_*)."""
import ctypes
import zlib

import numpy as np
import pytest

from openeat_b200 import _lib
from oracle import flac as oflac

# RFC 9639 D.1: two channels, one 16-bit sample each, verbatim subframes with 2 and 4 wasted bits
RFC_EXAMPLE_1 = bytes.fromhex(
    '664c614380000022100010000000' '0f00000f0ac442f000000001' '3e84b41807dc690307586a3dad1a2e0f'
    'fff869180000bf' '0358fd' '03128b' 'aa9a')
RFC_EXAMPLE_1_PCM = [[25588], [10416]]
# RFC 9639 D.3: mono, 8 bit, 32 kHz, 24 samples, one linear-predictor subframe (order 3, 4-bit precision, Rice k = 3)
RFC_EXAMPLE_3 = bytes.fromhex(
    '664c614380000022100010000000' '1f00001f07d0007000000018' 'f8f9e396f5cbcfc6dc807f9977906b32'
    'fff8680200' '17' 'e944004f6f313d1047d227cb6d090831452bdc28222280' '57a3')
RFC_EXAMPLE_3_PCM = [[0, 79, 111, 78, 8, -61, -90, -68, -13, 42, 67, 53, 13, -27, -46, -38, -12, 14, 24, 19, 6, -4, -5, 0]]


def _lib_decode(data, channel=0, first=0, count=None, verify=1):
    lib = _lib.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    sr, nch, bits, total = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
    rc = lib.oe_flac_info(buf.ctypes.data, buf.size, ctypes.byref(sr), ctypes.byref(nch), ctypes.byref(bits), ctypes.byref(total))
    if rc:
        raise ValueError(lib.oe_last_error().decode())
    if count is None:
        count = total.value - first
    out = np.full(max(count, 1), -12345, dtype=np.int32)
    n = ctypes.c_int64()
    rc = lib.oe_flac_decode(buf.ctypes.data, buf.size, channel, first, count, out.ctypes.data, verify, ctypes.byref(n))
    if rc:
        raise ValueError(lib.oe_last_error().decode())
    assert n.value == total.value
    return out[:count], dict(sample_rate=sr.value, channels=nch.value, bits=bits.value, total=total.value)


@pytest.mark.parametrize('data,pcm,rate,bits', [(RFC_EXAMPLE_1, RFC_EXAMPLE_1_PCM, 44100, 16),
                                                (RFC_EXAMPLE_3, RFC_EXAMPLE_3_PCM, 32000, 8)])
def test_rfc9639_worked_examples(data, pcm, rate, bits):
    """Oracle and product decoder against the RFC's own streams (both CRCs and the MD5 signature verified)."""
    got, info = oflac.decode(data, verify_md5=True)
    assert got.tolist() == pcm and info['sample_rate'] == rate and info['bits'] == bits
    for c, want in enumerate(pcm):
        out, li = _lib_decode(data, channel=c)
        assert out.tolist() == want
        assert li == dict(sample_rate=rate, channels=len(pcm), bits=bits, total=len(want))


def _signal(rng, nch, n, bits, kind):
    full = 1 << (bits - 1)
    if kind == 'noise':
        return rng.integers(-full, full, (nch, n))
    if kind == 'edges':                                      # the extreme codes, alternating: the widest residuals
        x = np.where(rng.integers(0, 2, (nch, n)) == 1, full - 1, -full)
        return x
    t = np.arange(n)
    x = np.stack([np.round(0.6 * full * np.sin(0.01 * (c + 1) * t + c)) for c in range(nch)]).astype(np.int64)
    return np.clip(x + rng.integers(-3, 4, (nch, n)), -full, full - 1)


CASES = []
for kind, order in (('constant', 0), ('verbatim', 0), ('fixed', 0), ('fixed', 1), ('fixed', 2), ('fixed', 3), ('fixed', 4),
                    ('lpc', 1), ('lpc', 8), ('lpc', 32)):
    CASES.append(dict(kind=kind, order=order))
CASES += [dict(kind='fixed', order=2, pbits=5), dict(kind='fixed', order=2, porder=3), dict(kind='lpc', order=4, porder=4, pbits=5),
          dict(kind='fixed', order=1, escape=True), dict(kind='lpc', order=6, escape=True, porder=2),
          dict(kind='fixed', order=2, stereo='left_side'), dict(kind='fixed', order=2, stereo='side_right'),
          dict(kind='lpc', order=5, stereo='mid_side'), dict(kind='verbatim', order=0, stereo='mid_side'),
          dict(kind='fixed', order=2, wasted=3), dict(kind='verbatim', order=0, wasted=5),
          dict(kind='fixed', order=2, variable=True, block_sizes=[17, 256, 1, 400, 95]),
          dict(kind='lpc', order=3, block=192), dict(kind='fixed', order=2, block=1000), dict(kind='fixed', order=2, block=4608),
          dict(kind='fixed', order=2, header_rate=True), dict(kind='fixed', order=2, header_bits=False),
          dict(kind='fixed', order=2, padding_block=40), dict(kind='fixed', order=3, with_md5=False)]


@pytest.mark.parametrize('case', CASES, ids=lambda c: '-'.join('%s' % v for v in c.values()))
@pytest.mark.parametrize('bits', [16, 8, 24])
def test_every_branch_of_the_format(case, bits):
    """Streams from the oracle's encoder, one per subframe type / predictor order / Rice parameter width / partition
    order / escape partition / stereo mode / wasted bits / block-size strategy: product decoder == input PCM == oracle
    decoder, MD5 verified by both."""
    if bits != 16 and case.get('order', 0) > 8:
        pytest.skip('one sample size is enough for the long predictors')
    rng = np.random.default_rng(zlib.crc32(str(sorted(case.items())).encode()) + bits)
    kw = dict(case)
    nch = 2 if kw.get('stereo', 'indep') != 'indep' or kw['kind'] == 'verbatim' else 1
    n = sum(kw['block_sizes']) if 'block_sizes' in kw else 769 + 2 * kw.get('block', 4096) // 4
    n = min(n, 2600)
    sig = 'noise' if kw['kind'] == 'verbatim' or kw.get('escape') else 'tone'
    pcm = _signal(rng, nch, n, bits, sig)
    if kw['kind'] == 'constant':
        pcm[:] = pcm[:, :1]
    if kw.get('wasted'):
        pcm = (pcm >> kw['wasted']) << kw['wasted']
    if kw['kind'] == 'lpc':
        kw['lpc'] = (8, 7, rng.integers(-100, 101, kw['order']).tolist())      # (precision, shift, coefficients)
    if kw.get('block') and kw['block'] > 4096:
        pcm = _signal(rng, nch, 4608 + 100, bits, sig)
        if kw.get('wasted'):
            pcm = (pcm >> kw['wasted']) << kw['wasted']
    data = oflac.encode(pcm, 16000, bits, **kw)
    ref, info = oflac.decode(data, verify_md5=True)
    assert np.array_equal(ref, pcm)
    for c in range(nch):
        out, li = _lib_decode(data, channel=c)
        assert np.array_equal(out, pcm[c])
        assert li['sample_rate'] == 16000 and li['channels'] == nch and li['bits'] == bits and li['total'] == pcm.shape[1]


def test_extreme_sample_values_and_32_bit_side_channel():
    """Full-scale alternating samples (the widest residuals; the side channel of a 16-bit stereo pair needs 17 bits) and
    24-bit full scale through an order-4 fixed predictor (the intermediate sums need more than 32 bits)."""
    rng = np.random.default_rng(7)
    for bits, stereo, kind, order in ((16, 'left_side', 'fixed', 4), (16, 'mid_side', 'fixed', 2), (24, 'side_right', 'fixed', 4),
                                      (24, 'indep', 'lpc', 12), (8, 'mid_side', 'verbatim', 0)):
        pcm = _signal(rng, 2, 700, bits, 'edges')
        lpc = (15, 15, [(1 << 14) - 1] * order) if kind == 'lpc' else None
        data = oflac.encode(pcm, 16000, bits, kind=kind, order=order, stereo=stereo, pbits=5, block=256, lpc=lpc)
        for c in range(2):
            out, _ = _lib_decode(data, channel=c)
            assert np.array_equal(out, pcm[c])


def test_segments_and_unannounced_length():
    rng = np.random.default_rng(3)
    pcm = _signal(rng, 1, 5000, 16, 'tone')
    data = bytearray(oflac.encode(pcm, 16000, 16, block=576, kind='fixed', order=2))
    for first, count in ((0, 5000), (0, 1), (575, 2), (576, 576), (1000, 3333), (4999, 1), (5000, 0)):
        out, _ = _lib_decode(bytes(data), first=first, count=count, verify=0)
        assert np.array_equal(out, pcm[0, first:first + count])
    # STREAMINFO without a sample count (a streamed encode): oe_flac_info counts by decoding
    body = 8 + 10
    data[body + 3] &= 0xF0
    data[body + 4:body + 8] = bytes(4)
    out, li = _lib_decode(bytes(data), verify=0)
    assert li['total'] == 5000 and np.array_equal(out, pcm[0])
    lib = _lib.load()
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    o = np.zeros(8, dtype=np.int32)
    assert lib.oe_flac_decode(buf.ctypes.data, buf.size, 0, 4999, 5, o.ctypes.data, 0, None) != 0
    assert 'requested' in lib.oe_last_error().decode()
    assert lib.oe_flac_decode(buf.ctypes.data, buf.size, 1, 0, 5, o.ctypes.data, 0, None) != 0        # no such channel


def test_corruption_is_detected():
    """Every single-byte corruption of the audio frames is caught (CRC-8 of the header, CRC-16 of the frame), truncation
    is caught, and a wrong signature is caught when verification is on -- nothing decodes to silently wrong samples."""
    rng = np.random.default_rng(5)
    pcm = _signal(rng, 2, 600, 16, 'tone')
    good = oflac.encode(pcm, 16000, 16, block=192, kind='fixed', order=2, stereo='mid_side')
    _, pos = oflac.parse_streaminfo(good)
    for at in list(range(pos, pos + 40)) + rng.integers(pos, len(good), 60).tolist():
        bad = bytearray(good)
        bad[at] ^= 1 << int(rng.integers(0, 8))
        with pytest.raises(ValueError):
            _lib_decode(bytes(bad), verify=0)
    for cut in (len(good) - 1, len(good) - 2, pos + 7, pos + 1, 30, 6):
        with pytest.raises(ValueError):
            _lib_decode(good[:cut], verify=0)
    bad = bytearray(good)
    bad[8 + 18 + 3] ^= 0x10                                       # the MD5 field of STREAMINFO
    assert np.array_equal(_lib_decode(bytes(bad), verify=0)[0], pcm[0])
    with pytest.raises(ValueError, match='MD5'):
        _lib_decode(bytes(bad), verify=1)
    with pytest.raises(ValueError):
        _lib_decode(b'fLaC' + bytes(64))


def test_read_wav_and_native_ingest_take_flac(tmp_path):
    """dataset.py:55-75 for .flac entries: read_wav (whole file and 'path,start,end' segments, 16-bit -> int16, 24-bit ->
    fp32 on the int16 scale) and the native ingest (a reader thread decodes the file into the pinned buffer) agree with
    the PCM that was encoded; a 24-bit stream is reported by the int16 ingest, not dropped silently."""
    from openeat_b200.dataset import read_wav
    from openeat_b200.ingest import NativeIngest
    rng = np.random.default_rng(11)
    a = _signal(rng, 1, 16000 + 37, 16, 'tone')
    b = _signal(rng, 2, 9000, 16, 'tone')
    c = _signal(rng, 1, 3000, 24, 'tone')
    (tmp_path / 'a.flac').write_bytes(oflac.encode(a, 16000, 16, block=4096, kind='lpc', order=8, porder=3, lpc=(12, 10, [1800, -900, 100, 20, -10, 5, 0, 3])))
    (tmp_path / 'b.flac').write_bytes(oflac.encode(b, 8000, 16, block=1152, kind='fixed', order=2, stereo='mid_side'))
    (tmp_path / 'c.flac').write_bytes(oflac.encode(c, 16000, 24, block=1024, kind='fixed', order=3))
    p = str(tmp_path)
    x, sr = read_wav(p + '/a.flac')
    assert sr == 16000 and x.dtype == np.int16 and np.array_equal(x, a[0])
    x, sr = read_wav(p + '/b.flac')
    assert sr == 8000 and np.array_equal(x, b[0])
    seg, _ = read_wav(p + '/a.flac', '0.25', '0.75')
    assert np.array_equal(seg, a[0, 4000:12000])
    tail, _ = read_wav(p + '/a.flac', '0.9', '5.0')
    assert np.array_equal(tail, a[0, 14400:])
    x, _ = read_wav(p + '/c.flac')
    assert x.dtype == np.float32 and np.array_equal(x, (c[0].astype(np.float32) / np.float32(1 << 23)) * np.float32(1 << 15))

    entries = [p + '/a.flac', p + '/b.flac', p + '/a.flac,0.25,0.75', p + '/c.flac', p + '/a.flac,0.9,5.0']
    ing = NativeIngest(threads=3, ring=2)
    for _ in range(3):
        buf, offs, lens, rates, loaded, slot = ing.load(entries, keys=['k%d' % i for i in range(5)])
        y = buf.numpy()
        assert loaded.tolist() == [True, True, True, False, True]
        assert rates.tolist()[:3] == [16000, 8000, 16000] and (offs % 8 == 0).all()
        assert lens.tolist() == [16037, 9000, 8000, 0, 1637]
        assert np.array_equal(y[offs[0]:offs[0] + lens[0]], a[0])
        assert np.array_equal(y[offs[1]:offs[1] + lens[1]], b[0])
        assert np.array_equal(y[offs[2]:offs[2] + lens[2]], a[0, 4000:12000])
        assert np.array_equal(y[offs[4]:offs[4] + lens[4]], a[0, 14400:])
    assert '24-bit FLAC' in ing.lib.oe_ingest_error(ing.handle, 3).decode()


# ---------------------------------------------------------------------------------------------------------------------
# GPU decoder (csrc/oe_flac_gpu.cuh): host side (encoder, frame index) and the kernel body emulated on the host; the real
# kernel runs in tests/test_gpu_mirrors.py (-m gpu) on the same streams.
# ---------------------------------------------------------------------------------------------------------------------
def lib_encode(pcm, rate=16000, block=4096, porder=3):
    lib = _lib.load()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    out = np.empty(2 * pcm.size + 8192, dtype=np.uint8)
    nbytes = ctypes.c_int64()
    assert lib.oe_flac_encode(pcm.ctypes.data, pcm.size, rate, block, porder, out.ctypes.data, out.size, ctypes.byref(nbytes)) == 0
    return out[:nbytes.value].tobytes()


def speechlike(rng, n, level=1500.0):
    """Low-pass noise under a syllable-rate envelope with pauses: its FLAC stream is ~0.53 of the PCM, LibriSpeech's own
    ratio (train-clean-100: 6.3 GB of FLAC for 100.6 h = 11.6 GB of PCM)."""
    x = rng.normal(0, 1, n + 64)
    for _ in range(3):
        x = np.convolve(x, [0.25, 0.5, 0.25], mode='same')
    x /= x.std()
    env = np.clip(np.sin(2 * np.pi * np.arange(n + 64) / 5000.0 + rng.uniform(0, 6)), 0.02, None)
    return np.clip(np.round(level * x * env)[:n], -32768, 32767).astype(np.int16)


def flac_stream_matrix(rng):
    """name -> (bytes, expected channel-0 int16 PCM): the product encoder's streams and the oracle encoder's (linear
    predictors, escapes, wasted bits, 4-bit parameters, variable block sizes ...), all mono 16 bit."""
    out = {}
    for name, n, block, po in (('enc4096', 20000, 4096, 3), ('enc1152', 5000, 1152, 5), ('enc_short', 37, 4096, 3),
                               ('enc_one', 1, 4096, 0), ('enc_odd', 4097, 4096, 2), ('enc256', 1000, 256, 8)):
        x = speechlike(rng, n)
        out[name] = (lib_encode(x, 16000, block, po), x)
    x = np.full(3000, -1234, dtype=np.int16)
    out['enc_constant'] = (lib_encode(x), x)
    x = rng.integers(-32768, 32768, 3000).astype(np.int16)
    out['enc_fullscale_noise'] = (lib_encode(x, block=576), x)
    for i, kw in enumerate([dict(kind='lpc', order=8, porder=3, lpc=(12, 10, [1800, -900, 100, 20, -10, 5, 0, 3])),
                            dict(kind='lpc', order=12, porder=2, pbits=5, lpc=(14, 12, [6000, -3000, 900, 20, -10, 5, 0, 3, 1, -1, 2, -2])),
                            dict(kind='lpc', order=1, lpc=(5, 3, [7])), dict(kind='fixed', order=4, porder=4),
                            dict(kind='fixed', order=0), dict(kind='fixed', order=1, escape=True, porder=2),
                            dict(kind='verbatim', order=0), dict(kind='fixed', order=2, wasted=3),
                            dict(kind='fixed', order=3, variable=True, block_sizes=[17, 256, 1, 400, 95, 2000]),
                            dict(kind='fixed', order=2, pbits=4, block=192), dict(kind='fixed', order=2, padding_block=40, header_rate=True),
                            dict(kind='fixed', order=4, porder=6, block=256)]):
        n = sum(kw['block_sizes']) if 'block_sizes' in kw else 2500
        x = speechlike(rng, n).astype(np.int64)
        if kw.get('wasted'):
            x = (x >> 3) << 3
        if kw['kind'] == 'verbatim' or kw.get('escape'):
            x = rng.integers(-32768, 32768, n)
        out['oracle%d' % i] = (oflac.encode(x[None], 16000, 16, **kw), x.astype(np.int16))
    return out


def test_product_encoder_streams_decode_with_the_oracle():
    """oe_flac_encode -> valid streams: the oracle decoder (pinned by the RFC's examples) checks both CRCs and the MD5
    signature and returns the PCM that went in; compression of the speech-like signal is reported by the bench."""
    rng = np.random.default_rng(21)
    for name, (data, pcm) in flac_stream_matrix(rng).items():
        if not name.startswith('enc'):
            continue
        got, info = oflac.decode(data, verify_md5=True)
        assert info['channels'] == 1 and info['bits'] == 16 and info['total'] == len(pcm), name
        assert np.array_equal(got[0], pcm), name
        out, _ = _lib_decode(data)
        assert np.array_equal(out, pcm), name
    x = speechlike(rng, 64000)
    assert 0.45 * 2 * len(x) < len(lib_encode(x)) < 0.6 * 2 * len(x)


def emul_decode(batch, verify=1):
    """Runs the kernel body of oe_flac_gpu.cuh on the host (liboe_emul.so) over a packed batch."""
    emul = ctypes.CDLL(_lib.EMUL_PATH)
    pcm = np.full(max(batch.total, 8), 77, dtype=np.int16)
    err = np.zeros(max(len(batch.lens), 1), dtype=np.int32)
    emul.oe_emul_flac_decode(ctypes.c_void_p(batch.comp.data_ptr()), ctypes.c_longlong(batch.comp_bytes),
                             ctypes.c_void_p(batch.frames.data_ptr()), ctypes.c_longlong(batch.n_frames),
                             ctypes.c_void_p(pcm.ctypes.data), ctypes.c_void_p(err.ctypes.data), ctypes.c_int(verify))
    return pcm, err


def test_frame_index_and_emulated_kernel(tmp_path):
    """oe_flac_pack + the kernel body (emulated thread per frame): every stream of the matrix, whole files and segments,
    equals the PCM that was encoded; alignment padding between utterances is never written; a second pack re-uses the ring."""
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
    for _ in range(3):
        b = ing.pack(entries)
        assert b.loaded.all() and (b.offsets % 8 == 0).all()
        assert b.lens.tolist() == [len(w) for w in want]
        pcm, err = emul_decode(b)
        assert not err.any()
        for i, w in enumerate(want):
            assert np.array_equal(pcm[b.offsets[i]:b.offsets[i] + b.lens[i]], w), entries[i]
            pad = pcm[b.offsets[i] + b.lens[i]:(b.offsets[i + 1] if i + 1 < len(want) else b.total)]
            assert (pad == 77).all()
    assert b.h2d_bytes < 2 * sum(len(w) for w in want[:len(streams)]) + 64 * len(entries) + 48 * b.n_frames


def test_emulated_kernel_flags_what_the_host_cannot_see(tmp_path):
    """Corruption inside a frame's body is invisible to the header walk: the kernel's end-of-frame / CRC-16 checks flag the
    entry (and only that entry).  Streams the GPU decoder does not take are reported by the pack, not dropped silently."""
    from openeat_b200.ingest import FlacGpuIngest
    rng = np.random.default_rng(23)
    x = speechlike(rng, 12000)
    good = lib_encode(x, block=1152)
    _, pos = oflac.parse_streaminfo(good)
    ing = FlacGpuIngest(threads=2, ring=2)
    (tmp_path / 'good.flac').write_bytes(good)
    flagged = 0
    for trial in range(40):
        bad = bytearray(good)
        at = int(rng.integers(pos + 8, len(good) - 2))
        bad[at] ^= 1 << int(rng.integers(0, 8))
        (tmp_path / 'bad.flac').write_bytes(bytes(bad))
        b = ing.pack([str(tmp_path / 'good.flac'), str(tmp_path / 'bad.flac')], report=False)
        if not b.loaded[1]:                                       # the flip hit a header: the host walk already refuses it
            continue
        pcm, err = emul_decode(b)
        assert err[0] == 0 and np.array_equal(pcm[b.offsets[0]:b.offsets[0] + b.lens[0]], x)
        assert err[1] != 0
        flagged += 1
    assert flagged >= 30
    # order-32 predictor: legal FLAC outside the kernel's range -> OE_FLAC_ERR_HOST; stereo / 24 bit: refused by the pack
    y = speechlike(rng, 2000).astype(np.int64)
    (tmp_path / 'o32.flac').write_bytes(oflac.encode(y[None], 16000, 16, kind='lpc', order=32, lpc=(8, 7, rng.integers(-50, 51, 32).tolist())))
    (tmp_path / 'st.flac').write_bytes(oflac.encode(np.stack([y, y]), 16000, 16, kind='fixed', order=2))
    (tmp_path / 'b24.flac').write_bytes(oflac.encode(y[None] * 200, 16000, 24, kind='fixed', order=2))
    (tmp_path / 'b8.flac').write_bytes(oflac.encode(y[None] >> 6, 16000, 8, kind='fixed', order=1))
    b8 = ing.pack([str(tmp_path / 'b8.flac')], report=False)
    assert not b8.loaded[0] and '8-bit' in ing.lib.oe_ingest_error(ing.handle, 0).decode()
    b = ing.pack([str(tmp_path / n) for n in ('o32.flac', 'st.flac', 'b24.flac', 'good.flac', 'missing.flac')], report=False)
    assert b.loaded.tolist() == [True, False, False, True, False]
    errs = [ing.lib.oe_ingest_error(ing.handle, i).decode() for i in range(5)]
    assert 'channels' in errs[1] and '24-bit' in errs[2] and 'No such file' in errs[4]
    pcm, err = emul_decode(b)
    assert err[0] == 4 and err[3] == 0 and np.array_equal(pcm[b.offsets[3]:b.offsets[3] + b.lens[3]], x)


def test_pack_ahead_generator_is_race_free(tmp_path):
    """flac_gpu_batches: batches packed ahead by two workers (one handle each, one pack at a time per handle) arrive in
    order and decode (emulated kernel) to the PCM that was encoded, over many small batches."""
    from openeat_b200.ingest import flac_gpu_batches
    rng = np.random.default_rng(41)
    pcm, items = {}, []
    for i in range(12):
        x = speechlike(rng, int(rng.integers(300, 9000)))
        p = tmp_path / ('f%d.flac' % i)
        p.write_bytes(lib_encode(x, block=576))
        pcm[str(p)] = x
        items.append(('k%d' % i, str(p), [1], 1.0))
    batches = [[items[j] for j in rng.permutation(12)[:int(rng.integers(1, 8))]] for _ in range(150)]
    seen = 0
    for want, got in zip(batches, flac_gpu_batches(iter(batches), depth=3, workers=2, threads=4)):
        b = got[0]
        assert got[3] == [x[0] for x in want] and b.loaded.all()
        out, err = emul_decode(b)
        assert not err.any()
        for i, it in enumerate(want):
            assert np.array_equal(out[b.offsets[i]:b.offsets[i] + b.lens[i]], pcm[it[1]])
        got[8](None)
        seen += 1
    assert seen == 150


def test_async_pack_reports_bad_entries_and_grows_its_buffers(tmp_path, capsys):
    """oe_flac_submit / oe_flac_wait: a first batch larger than the initial ring slot is re-submitted with larger buffers;
    unreadable entries come back not loaded with their reason printed (dataset.py:108-111), the others decode."""
    from openeat_b200.ingest import FlacGpuIngest
    rng = np.random.default_rng(51)
    x = speechlike(rng, 900000)                                    # ~0.9 MB of FLAC: with the others, more than the 1 MB first slot
    y = speechlike(rng, 5000)
    (tmp_path / 'big.flac').write_bytes(lib_encode(x))
    (tmp_path / 'big2.flac').write_bytes(lib_encode(x[::-1].copy()))
    (tmp_path / 'small.flac').write_bytes(lib_encode(y))
    (tmp_path / 'bad.flac').write_bytes(b'fLaC' + bytes(64))
    names = [str(tmp_path / n) for n in ('big.flac', 'bad.flac', 'small.flac', 'missing.flac', 'big2.flac')]
    ing = FlacGpuIngest(threads=2, ring=3)
    tickets = [ing.submit(names, keys=list('abcde')) for _ in range(2)]
    for t in tickets:
        b = ing.wait(t)
        assert b.loaded.tolist() == [True, False, True, False, True] and b.lens.tolist() == [900000, 0, 5000, 0, 900000]
        out, err = emul_decode(b)
        assert not err.any()
        assert np.array_equal(out[b.offsets[0]:b.offsets[0] + 900000], x) and np.array_equal(out[b.offsets[2]:b.offsets[2] + 5000], y)
        assert np.array_equal(out[b.offsets[4]:b.offsets[4] + 900000], x[::-1])
    text = capsys.readouterr().out
    assert 'STREAMINFO' in text and 'No such file' in text


def test_shard_tar_with_flac_members(tmp_path):
    """wenet-style shards (data_type='shard'): <key>.flac members are decoded like <key>.wav ones."""
    import io
    import tarfile
    from openeat_b200.processor import tar_file_and_group
    rng = np.random.default_rng(61)
    x, y = speechlike(rng, 7000), speechlike(rng, 3000)
    shard = tmp_path / 's0.tar'
    with tarfile.open(shard, 'w') as tar:
        for name, blob in (('u1.flac', lib_encode(x, block=1152)), ('u1.txt', 'hello world'.encode()),
                           ('u2.txt', 'b'.encode()), ('u2.flac', lib_encode(y))):
            info = tarfile.TarInfo(name)
            info.size = len(blob)
            tar.addfile(info, io.BytesIO(blob))
    got = list(tar_file_and_group([str(shard)]))
    assert [s['key'] for s in got] == ['u1', 'u2'] and [s['txt'] for s in got] == ['hello world', 'b']
    assert np.array_equal(got[0]['wav'], x) and np.array_equal(got[1]['wav'], y) and got[0]['sample_rate'] == 16000


def test_id3v2_tag_in_front_of_the_stream(tmp_path):
    """Taggers sometimes prepend an ID3v2 tag to a .flac file; libFLAC (behind torchaudio.load) skips it, so do read_wav,
    the native ingest and the GPU pack (frame offsets are absolute: the kernel never sees the tag)."""
    from openeat_b200.dataset import read_wav
    from openeat_b200.ingest import FlacGpuIngest, NativeIngest
    rng = np.random.default_rng(71)
    x = speechlike(rng, 9000)
    size = 301                                                     # tag body, sync-safe size field; 0xFF bytes inside must not fool the frame walk
    tag = b'ID3\x04\x00\x00' + bytes([(size >> 21) & 127, (size >> 14) & 127, (size >> 7) & 127, size & 127]) + b'\xff\xf8' * 150 + b'\x00'
    assert len(tag) == 10 + size
    p = tmp_path / 'tagged.flac'
    p.write_bytes(tag + lib_encode(x, block=1152))
    got, sr = read_wav(str(p))
    assert sr == 16000 and np.array_equal(got, x)
    seg, _ = read_wav(str(p), '0.1', '0.3')
    assert np.array_equal(seg, x[1600:4800])
    buf, offs, lens, rates, loaded, _ = NativeIngest(threads=2, ring=2).load([str(p)])
    assert loaded.all() and np.array_equal(buf.numpy()[offs[0]:offs[0] + lens[0]], x)
    b = FlacGpuIngest(threads=2, ring=2).pack([str(p), str(p) + ',0.1,0.3'])
    assert b.loaded.all()
    out, err = emul_decode(b)
    assert not err.any() and np.array_equal(out[b.offsets[0]:b.offsets[0] + 9000], x)
    assert np.array_equal(out[b.offsets[1]:b.offsets[1] + b.lens[1]], x[1600:4800])
