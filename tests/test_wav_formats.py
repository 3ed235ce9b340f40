"""read_wav (openeat/dataset/dataset.py:62-75: torchaudio.load + ``* (1 << 15)``) on the WAV encodings torchaudio's
backends read, checked against scipy.io.wavfile as an independent decoder and against the normalisation torchaudio
documents (int / 2^(bits-1), 8-bit unsigned offset 128, float as is)."""
import os
import struct

import numpy as np
import pytest

from openeat_b200.dataset import read_wav


def _write(path, tag, bits, nch, sr, frames, extensible=False, junk=True, streamed=False):
    """frames: (n, nch) array already in the file's sample domain."""
    if tag == 3:
        raw = frames.astype('<f4' if bits == 32 else '<f8').tobytes()
    elif bits == 8:
        raw = frames.astype(np.uint8).tobytes()
    elif bits == 16:
        raw = frames.astype('<i2').tobytes()
    elif bits == 24:
        v = frames.astype(np.int64) & 0xFFFFFF
        raw = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255], axis=-1).astype(np.uint8).tobytes()
    else:
        raw = frames.astype('<i4').tobytes()
    block = bits // 8 * nch
    if extensible:
        guid = struct.pack('<H', tag) + b'\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71'
        fmt = struct.pack('<HHIIHHHHI', 0xFFFE, nch, sr, sr * block, block, bits, 22, bits, 0) + guid
    else:
        fmt = struct.pack('<HHIIHH', tag, nch, sr, sr * block, block, bits)
    body = b'WAVE' + b'fmt ' + struct.pack('<I', len(fmt)) + fmt
    if junk:
        body += b'LIST' + struct.pack('<I', 5) + b'abcde' + b'\x00'          # odd-sized chunk + pad byte
    body += b'data' + struct.pack('<I', 0xFFFFFFFF if streamed else len(raw)) + raw
    with open(path, 'wb') as f:
        f.write(b'RIFF' + struct.pack('<I', len(body)) + body)


CASES = [(1, 8), (1, 16), (1, 24), (1, 32), (3, 32), (3, 64)]


@pytest.mark.parametrize('tag,bits', CASES)
@pytest.mark.parametrize('nch', [1, 2])
def test_read_wav_matches_scipy_and_torchaudio_scaling(tmp_path, tag, bits, nch):
    from scipy.io import wavfile
    rng = np.random.default_rng(bits + nch)
    n = 1000
    if tag == 3:
        frames = rng.uniform(-1, 1, (n, nch))
    elif bits == 8:
        frames = rng.integers(0, 256, (n, nch))
    else:
        frames = rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), (n, nch))
    p = str(tmp_path / 'a.wav')
    _write(p, tag, bits, nch, 22050, frames)
    got, sr = read_wav(p)
    assert sr == 22050 and got.shape == (n,)
    ref_sr, ref = wavfile.read(p)
    ref = ref.reshape(n, nch)[:, 0]
    assert ref_sr == sr
    if tag == 3:
        want = ref.astype(np.float32) * np.float32(32768)
    elif bits == 8:
        want = (ref.astype(np.float32) - 128.0) / 128.0 * np.float32(32768)
    elif bits == 16:
        assert got.dtype == np.int16
        want = ref
    elif bits == 24:
        # scipy returns 24-bit samples left-justified in int32
        want = (ref.astype(np.int64) >> 8).astype(np.float32) / np.float32(1 << 23) * np.float32(32768)
    else:
        want = ref.astype(np.float32) / np.float32(2.0 ** 31) * np.float32(32768)
    assert np.array_equal(np.asarray(got), want)
    # the file's own values, independent of scipy
    src = frames[:, 0]
    if tag == 1 and bits == 24:
        assert np.array_equal(got, src.astype(np.float32) / np.float32(256))          # s / 2^23 * 2^15, exact in fp32


def test_read_wav_extensible_streamed_and_segment(tmp_path):
    rng = np.random.default_rng(0)
    frames = rng.integers(-30000, 30000, (16000, 1))
    p = str(tmp_path / 'e.wav')
    _write(p, 1, 16, 1, 16000, frames, extensible=True, streamed=True)
    got, sr = read_wav(p)
    assert sr == 16000 and np.array_equal(got, frames[:, 0])
    seg, _ = read_wav(p, '0.25', '0.5')              # frame_offset = int(0.25 * sr), num_frames = int(0.5 * sr) - offset
    assert np.array_equal(seg, frames[4000:8000, 0])
    tail, _ = read_wav(p, '0.9', '1.5')              # clipped at the end of the file
    assert np.array_equal(tail, frames[14400:, 0])


def test_read_wav_rejects_what_it_cannot_decode(tmp_path):
    p = str(tmp_path / 'x.flac')
    with open(p, 'wb') as f:
        f.write(b'fLaC' + bytes(64))
    with pytest.raises(ValueError, match='STREAMINFO'):           # a damaged FLAC stream (sound ones: tests/test_flac.py)
        read_wav(p)
    q = str(tmp_path / 'adpcm.wav')
    _write(q, 2, 4, 1, 8000, np.zeros((8, 1)), junk=False)       # MS ADPCM tag
    with pytest.raises(ValueError, match='format tag 2'):
        read_wav(q)
