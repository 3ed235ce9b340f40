"""FLAC (RFC 9639) decoder and a small encoder, pure Python -- TEST INFRASTRUCTURE ONLY.

The reference loads audio with ``torchaudio.load`` (``openeat/dataset/dataset.py:64-72``), i.e. whatever libsox reads;
the LibriSpeech recipe's corpus is FLAC.  libFLAC / libsox are third-party code absent from ``/root/reference`` and
from this image, so the oracle restates the published format (RFC 9639, "Free Lossless Audio Codec"):

* ``decode(data)``  -- the decoder of section 9 ("Frame structure") and 9.2 ("Subframes"): constant / verbatim / fixed /
  linear-predictor subframes, partitioned Rice residuals with 4- and 5-bit parameters and escape partitions, wasted
  bits, the three stereo decorrelation modes, CRC-8 of the frame header, CRC-16 of the frame, MD5 of the decoded PCM.
  Pinned by the RFC's own worked examples (Appendix D; ``tests/test_flac.py`` holds their bytes, both CRCs and the
  MD5 signature check out), which is the only golden material there is without an encoder in the image.
* ``encode(...)``   -- a fixture generator: it can be told which subframe type, predictor order, Rice parameter width,
  partition order, block size and stereo mode to use, so that the product decoder (``csrc/oe_flac.h``) meets every
  branch of the format, not only the ones a particular encoder would choose.

Pure-Python loops: small cases only.
"""
import hashlib
import struct

import numpy as np


def crc8(data):
    c = 0
    for x in data:
        c ^= x
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def crc16(data):
    c = 0
    for x in data:
        c ^= x << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


class FlacError(ValueError):
    pass


# ------------------------------------------------------------------ decoder
class _Bits:
    def __init__(self, data, pos=0):
        self.d, self.p = data, pos * 8

    def u(self, n):
        v = 0
        for _ in range(n):
            byte = self.d[self.p >> 3]
            v = (v << 1) | ((byte >> (7 - (self.p & 7))) & 1)
            self.p += 1
        return v

    def s(self, n):
        if n == 0:
            return 0
        v = self.u(n)
        return v - (1 << n) if v >> (n - 1) else v

    def unary(self):
        q = 0
        while self.u(1) == 0:
            q += 1
        return q

    def align(self):
        self.p = (self.p + 7) & ~7

    @property
    def byte(self):
        return self.p >> 3


_BLOCK = {1: 192, 2: 576, 3: 1152, 4: 2304, 5: 4608, 8: 256, 9: 512, 10: 1024, 11: 2048, 12: 4096, 13: 8192,
          14: 16384, 15: 32768}
_RATE = {1: 88200, 2: 176400, 3: 192000, 4: 8000, 5: 16000, 6: 22050, 7: 24000, 8: 32000, 9: 44100, 10: 48000,
         11: 96000}
_BPS = {1: 8, 2: 12, 4: 16, 5: 20, 6: 24, 7: 32}
_FIXED = {0: (), 1: (1,), 2: (2, -1), 3: (3, -3, 1), 4: (4, -6, 4, -1)}


def parse_streaminfo(data):
    if data[:4] != b'fLaC':
        raise FlacError('not a FLAC stream')
    pos, info = 4, None
    while True:
        last, kind = data[pos] >> 7, data[pos] & 0x7F
        size = int.from_bytes(data[pos + 1:pos + 4], 'big')
        body = data[pos + 4:pos + 4 + size]
        if kind == 0:
            b = _Bits(body)
            info = dict(min_block=b.u(16), max_block=b.u(16), min_frame=b.u(24), max_frame=b.u(24), sample_rate=b.u(20),
                        channels=b.u(3) + 1, bits=b.u(5) + 1, total=b.u(36), md5=bytes(body[18:34]))
        pos += 4 + size
        if last:
            break
    if info is None:
        raise FlacError('no STREAMINFO block')
    return info, pos


def _residual(b, n, order, out):
    method = b.u(2)
    if method > 1:
        raise FlacError('reserved residual coding method')
    pbits = 4 if method == 0 else 5
    porder = b.u(4)
    if (n >> porder) << porder != n and porder:
        raise FlacError('block size not divisible by the partition count')
    for part in range(1 << porder):
        cnt = (n >> porder) - (order if part == 0 else 0)
        if cnt < 0:
            raise FlacError('partition shorter than the predictor order')
        k = b.u(pbits)
        if k == (1 << pbits) - 1:
            raw = b.u(5)
            for _ in range(cnt):
                out.append(b.s(raw))
        else:
            for _ in range(cnt):
                v = (b.unary() << k) | b.u(k)
                out.append((v >> 1) ^ -(v & 1))


def _subframe(b, n, bps):
    if b.u(1):
        raise FlacError('subframe padding bit set')
    kind = b.u(6)
    wasted = 0
    if b.u(1):
        wasted = b.unary() + 1
        bps -= wasted
    if kind == 0:
        s = [b.s(bps)] * n
    elif kind == 1:
        s = [b.s(bps) for _ in range(n)]
    elif 8 <= kind <= 12 or kind >= 32:
        if kind >= 32:
            order = kind - 31
            s = [b.s(bps) for _ in range(order)]
            prec = b.u(4) + 1
            if prec == 16:
                raise FlacError('reserved predictor precision')
            shift = b.s(5)
            if shift < 0:
                raise FlacError('negative predictor shift')
            coefs = [b.s(prec) for _ in range(order)]
        else:
            order = kind - 8
            s = [b.s(bps) for _ in range(order)]
            shift, coefs = 0, _FIXED[order]
        res = []
        _residual(b, n, order, res)
        for r in res:
            acc = 0
            for j, c in enumerate(coefs):
                acc += c * s[-1 - j]
            s.append(r + (acc >> shift))
    else:
        raise FlacError('reserved subframe type %d' % kind)
    return [x << wasted for x in s] if wasted else s


def decode(data, verify_md5=True):
    """-> (int32 array (channels, samples), info dict); every CRC is checked, the MD5 when the stream carries one."""
    data = bytes(data)
    info, pos = parse_streaminfo(data)
    chans = [[] for _ in range(info['channels'])]
    while pos < len(data):
        start = pos
        b = _Bits(data, pos)
        if b.u(15) != 0x7FFC:
            raise FlacError('lost frame sync at byte %d' % pos)
        b.u(1)
        bs_code, sr_code, ch_code, bps_code = b.u(4), b.u(4), b.u(4), b.u(3)
        if b.u(1):
            raise FlacError('reserved header bit set')
        first = b.u(8)                                    # UTF-8-like coded frame / sample number
        extra = 0
        while first & (0x80 >> extra):
            extra += 1
        for _ in range(max(0, extra - 1)):
            if b.u(8) >> 6 != 2:
                raise FlacError('bad coded number')
        if bs_code == 0:
            raise FlacError('reserved block size code')
        n = b.u(8) + 1 if bs_code == 6 else b.u(16) + 1 if bs_code == 7 else _BLOCK[bs_code]
        if sr_code == 12:
            b.u(8)
        elif sr_code in (13, 14):
            b.u(16)
        elif sr_code == 15:
            raise FlacError('invalid sample rate code')
        if crc8(data[start:b.byte]) != b.u(8):
            raise FlacError('frame header CRC-8 mismatch')
        bps = info['bits'] if bps_code == 0 else _BPS.get(bps_code)
        if bps is None:
            raise FlacError('reserved sample size code')
        if ch_code < 8:
            nch, subs = ch_code + 1, [_subframe(b, n, bps) for _ in range(ch_code + 1)]
        elif ch_code == 8:                                # left / side
            left, side = _subframe(b, n, bps), _subframe(b, n, bps + 1)
            nch, subs = 2, [left, [l - s for l, s in zip(left, side)]]
        elif ch_code == 9:                                # side / right
            side, right = _subframe(b, n, bps + 1), _subframe(b, n, bps)
            nch, subs = 2, [[s + r for s, r in zip(side, right)], right]
        elif ch_code == 10:                               # mid / side
            mid, side = _subframe(b, n, bps), _subframe(b, n, bps + 1)
            full = [(m << 1) | (s & 1) for m, s in zip(mid, side)]
            nch, subs = 2, [[(m + s) >> 1 for m, s in zip(full, side)], [(m - s) >> 1 for m, s in zip(full, side)]]
        else:
            raise FlacError('reserved channel assignment')
        if nch != info['channels']:
            raise FlacError('channel count changes inside the stream')
        b.align()
        if crc16(data[start:b.byte]) != b.u(16):
            raise FlacError('frame CRC-16 mismatch')
        for c in range(nch):
            chans[c].extend(subs[c])
        pos = b.byte
    pcm = np.array(chans, dtype=np.int64).reshape(info['channels'], -1)
    if info['total'] and pcm.shape[1] != info['total']:
        raise FlacError('STREAMINFO announces %d samples, the frames hold %d' % (info['total'], pcm.shape[1]))
    if verify_md5 and any(info['md5']):
        if pcm_md5(pcm, info['bits']) != info['md5']:
            raise FlacError('MD5 signature mismatch')
    return pcm.astype(np.int32), info


def pcm_md5(pcm, bits):
    """MD5 of the interleaved little-endian samples, each in ceil(bits / 8) bytes (RFC 9639 section 8.2)."""
    nbytes = (bits + 7) // 8
    inter = np.ascontiguousarray(np.asarray(pcm, dtype=np.int64).T).reshape(-1)
    raw = inter.astype('<i8').view(np.uint8).reshape(-1, 8)[:, :nbytes]
    return hashlib.md5(raw.tobytes()).digest()


# ------------------------------------------------------------------ encoder (fixture generator)
class _Writer:
    def __init__(self):
        self.bits = []

    def u(self, v, n):
        assert 0 <= v < (1 << n) or n == 0, (v, n)
        self.bits.extend((v >> (n - 1 - i)) & 1 for i in range(n))

    def s(self, v, n):
        assert -(1 << (n - 1)) <= v < (1 << (n - 1)), (v, n)
        self.u(v & ((1 << n) - 1), n)

    def unary(self, q):
        self.bits.extend([0] * q + [1])

    def align(self):
        self.bits.extend([0] * (-len(self.bits) % 8))

    def bytes(self):
        assert len(self.bits) % 8 == 0
        return bytes(int(''.join(map(str, self.bits[i:i + 8])), 2) for i in range(0, len(self.bits), 8))


def _coded_number(v):
    if v < 0x80:
        return bytes([v])
    n = 2
    while v >= 1 << (5 * n + 1):
        n += 1
    out = [((0xFF << (8 - n)) & 0xFF) | (v >> (6 * (n - 1)))]
    for i in range(n - 2, -1, -1):
        out.append(0x80 | ((v >> (6 * i)) & 0x3F))
    return bytes(out)


def _write_residual(w, res, n, order, pbits, porder, escape):
    w.u(0 if pbits == 4 else 1, 2)
    w.u(porder, 4)
    at = 0
    for part in range(1 << porder):
        cnt = (n >> porder) - (order if part == 0 else 0)
        chunk = res[at:at + cnt]
        at += cnt
        if escape:
            raw = max([0] + [(abs(v) if v >= 0 else abs(v + 1)).bit_length() + 1 for v in chunk]) if any(chunk) else 0
            w.u((1 << pbits) - 1, pbits)
            w.u(raw, 5)
            for v in chunk:
                if raw:
                    w.s(v, raw)
        else:
            zz = [(v << 1) if v >= 0 else ((-v) << 1) - 1 for v in chunk]
            mean = (sum(zz) / len(zz)) if zz else 0
            k = min(max(int(mean).bit_length() - 1, 0), (1 << pbits) - 2)
            w.u(k, pbits)
            for z in zz:
                w.unary(z >> k)
                w.u(z & ((1 << k) - 1), k)


def _write_subframe(w, s, bps, kind, order, pbits, porder, escape, wasted, lpc):
    s = [int(v) for v in s]
    n = len(s)
    if wasted:
        assert all(v % (1 << wasted) == 0 for v in s)
        s = [v >> wasted for v in s]
        bps -= wasted
    w.u(0, 1)
    code = {'constant': 0, 'verbatim': 1, 'fixed': 8 + order, 'lpc': 31 + order}[kind]
    w.u(code, 6)
    w.u(1 if wasted else 0, 1)
    if wasted:
        w.unary(wasted - 1)
    if kind == 'constant':
        assert len(set(s)) == 1
        w.s(s[0], bps)
    elif kind == 'verbatim':
        for v in s:
            w.s(v, bps)
    else:
        for v in s[:order]:
            w.s(v, bps)
        if kind == 'lpc':
            prec, shift, coefs = lpc
            assert len(coefs) == order
            w.u(prec - 1, 4)
            w.s(shift, 5)
            for c in coefs:
                w.s(c, prec)
        else:
            shift, coefs = 0, _FIXED[order]
        res = []
        for i in range(order, n):
            acc = sum(c * s[i - 1 - j] for j, c in enumerate(coefs))
            res.append(s[i] - (acc >> shift))
        _write_residual(w, res, n, order, pbits, porder, escape)


def encode(pcm, sample_rate, bits, block=4096, kind='fixed', order=2, pbits=4, porder=0, escape=False, stereo='indep',
           wasted=0, lpc=None, variable=False, block_sizes=None, with_md5=True, padding_block=0, header_rate=False,
           header_bits=True):
    """pcm: (channels, samples) integers.  One configuration for every subframe of the stream (block_sizes: explicit
    list of block lengths, implies the variable-block-size strategy when variable=True)."""
    pcm = np.atleast_2d(np.asarray(pcm, dtype=np.int64))
    nch, total = pcm.shape
    out = bytearray(b'fLaC')
    sizes = list(block_sizes) if block_sizes else [min(block, total - i) for i in range(0, total, block)]
    assert sum(sizes) == total
    si = _Writer()
    si.u(min(sizes[:-1] or sizes), 16)
    si.u(max(sizes), 16)
    si.u(0, 24)
    si.u(0, 24)
    si.u(sample_rate, 20)
    si.u(nch - 1, 3)
    si.u(bits - 1, 5)
    si.u(total, 36)
    body = si.bytes() + (pcm_md5(pcm, bits) if with_md5 else bytes(16))
    out += bytes([0x00 if padding_block else 0x80]) + len(body).to_bytes(3, 'big') + body
    if padding_block:
        out += bytes([0x81]) + padding_block.to_bytes(3, 'big') + bytes(padding_block)
    rate_codes = {v: k for k, v in _RATE.items()}
    bps_codes = {v: k for k, v in _BPS.items()}
    at = 0
    for fi, n in enumerate(sizes):
        w = _Writer()
        w.u(0x7FFC, 15)
        w.u(1 if variable else 0, 1)
        bs_code = next((k for k, v in _BLOCK.items() if v == n), 6 if n <= 256 else 7)
        w.u(bs_code, 4)
        sr_code = rate_codes.get(sample_rate, 0) if header_rate else 0
        if header_rate and sr_code == 0:
            sr_code = 13
        w.u(sr_code, 4)
        w.u({'indep': nch - 1, 'left_side': 8, 'side_right': 9, 'mid_side': 10}[stereo], 4)
        w.u(bps_codes.get(bits, 0) if header_bits else 0, 3)
        w.u(0, 1)
        for byte in _coded_number(at if variable else fi):
            w.u(byte, 8)
        if bs_code == 6:
            w.u(n - 1, 8)
        elif bs_code == 7:
            w.u(n - 1, 16)
        if sr_code == 13:
            w.u(sample_rate, 16)
        w.u(crc8(w.bytes()), 8)
        blk = pcm[:, at:at + n]
        if stereo == 'indep':
            subs = [(blk[c], bits) for c in range(nch)]
        elif stereo == 'left_side':
            subs = [(blk[0], bits), (blk[0] - blk[1], bits + 1)]
        elif stereo == 'side_right':
            subs = [(blk[0] - blk[1], bits + 1), (blk[1], bits)]
        else:
            subs = [((blk[0] + blk[1]) >> 1, bits), (blk[0] - blk[1], bits + 1)]
        for s, b in subs:
            k = kind
            if k == 'constant' and len(set(int(v) for v in s)) != 1:
                k = 'verbatim'
            if k in ('fixed', 'lpc') and n <= order:
                k = 'verbatim'
            o = order if k in ('fixed', 'lpc') else 0
            po = porder
            while po and ((n >> po) << po != n or (n >> po) < o):
                po -= 1
            ws = wasted if all(int(v) % (1 << wasted) == 0 for v in s) else 0
            _write_subframe(w, s, b, k, o, pbits, po, escape, ws, lpc)
        w.align()
        frame = w.bytes()
        out += frame + crc16(frame).to_bytes(2, 'big')
        at += n
    return bytes(out)
