"""Accuracy side of the tensor-core question ("why not tensor cores"): a two-stage 16 x 16
DFT-as-GEMM of the packed 256-point complex FFT, with the operands split into bf16 / tf32 pieces the way a tcgen05
kernel would have to feed them (kind::f16 takes bf16, kind::tf32 takes 10-bit mantissas; accumulation is fp32),
emulated in numpy.  Everything outside the two GEMMs (window, twiddle, untangle, power, mel, log) runs in float64, so
the error reported is the DFT stages' alone.  Compared with the fp64 oracle on the golden signal classes.

    python tools/ubench/dft_tc_accuracy.py            # prints a table, writes profiles/r02_dft_tc_accuracy.json

CPU only (test infrastructure, like oracle/): never imported by the product.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fbank as F          # noqa: E402
from oracle import signals             # noqa: E402


def round_mantissa(x, bits):
    """fp32 -> `bits` explicit mantissa bits (round to nearest even), returned as fp32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    u = u + ((1 << (drop - 1)) - 1) + ((u >> drop) & 1)
    u = (u >> drop) << drop
    return u.astype(np.uint32).view(np.float32)


def split(x, bits, parts):
    out, r = [], np.asarray(x, dtype=np.float32)
    for _ in range(parts):
        p = round_mantissa(r, bits)
        out.append(p)
        r = (r - p).astype(np.float32)
    return out


def mm(a, b, bits, terms):
    """sum of a_i @ b_j over `terms` (fp32 accumulate); bits None = plain fp32 product."""
    if bits is None:
        return a.astype(np.float32) @ b.astype(np.float32)
    n = 1 + max(max(t) for t in terms)
    sa, sb = split(a, bits, n), split(b, bits, n)
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for i, j in sorted(terms, key=lambda t: -(t[0] + t[1])):          # small terms first
        acc = acc + sa[i] @ sb[j]
    return acc


def dft16_matrix():
    """real 32 x 32 matrix of the 16-point complex DFT acting on [re(16) | im(16)] row vectors."""
    k = np.arange(16)
    w = np.exp(-2j * np.pi * np.outer(k, k) / 16.0)       # [n, k]
    m = np.zeros((32, 32))
    m[:16, :16], m[:16, 16:] = w.real, w.imag
    m[16:, :16], m[16:, 16:] = -w.imag, w.real
    return m


def logmel_two_stage(wave, bits, terms, mel64):
    h = F.windowed_frames(wave, np.float32)                          # (m, 512) fp32, what the kernel's front produces
    m = h.shape[0]
    z = h[:, 0::2].astype(np.float64) + 1j * h[:, 1::2]              # (m, 256), n = 16 n1 + n2
    z = z.reshape(m, 16, 16)                                         # [frame, n1, n2]
    d = dft16_matrix()
    a = np.concatenate([z.real.transpose(0, 2, 1), z.imag.transpose(0, 2, 1)], axis=2).reshape(m * 16, 32)  # rows (frame, n2)
    y = mm(a, d, bits, terms).astype(np.float64).reshape(m, 16, 32)  # [frame, n2, (re k1 | im k1)]
    y = y[:, :, :16] + 1j * y[:, :, 16:]                             # [frame, n2, k1]
    tw = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 256.0)     # [n2, k1]
    y = (y * tw[None]).astype(np.complex64)                          # twiddled, rounded to fp32 like registers would be
    a2 = np.concatenate([y.real.transpose(0, 2, 1), y.imag.transpose(0, 2, 1)], axis=2).reshape(m * 16, 32)  # rows (frame, k1), cols n2
    zz = mm(a2, d, bits, terms).astype(np.float64).reshape(m, 16, 32)
    zz = zz[:, :, :16] + 1j * zz[:, :, 16:]                          # [frame, k1, k2] -> Z[k1 + 16 k2]
    Z = zz.transpose(0, 2, 1).reshape(m, 256)
    k = np.arange(257)
    Zk, Zc = Z[:, k % 256], np.conj(Z[:, (256 - k) % 256])
    X = 0.5 * (Zk + Zc) - 0.5j * np.exp(-2j * np.pi * k / 512.0) * (Zk - Zc)
    power = np.abs(X) ** 2
    melp = np.concatenate([mel64, np.zeros((mel64.shape[0], 1))], axis=1)
    return np.log(np.maximum(power @ melp.T, float(F.EPS_F32)))


VARIANTS = [
    ('fp32 operands (sanity)', None, None, 1.0),
    ('bf16 x1', 7, [(0, 0)], 1.0),
    ('bf16 split 2, 3 products', 7, [(0, 0), (0, 1), (1, 0)], 3.0),
    ('bf16 split 3, 6 products', 7, [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)], 6.0),
    ('tf32 x1', 10, [(0, 0)], 2.0),
    ('tf32 split 2, 3 products', 10, [(0, 0), (0, 1), (1, 0)], 6.0),
]


def main():
    mel64 = F.mel_banks(80, 512, 16000.0, 20.0, 0.0, np.float32).astype(np.float64)
    rows = []
    for name, bits, terms, cost in VARIANTS:
        row = {'variant': name, 'bf16_equivalent_passes': cost, 'max_abs_logmel_err': {}}
        for kind in ('white', 'speech', 'lsb', 'dcsine', 'square'):
            w = signals.make(kind, 8000, 7)
            ref = F.fbank(w.astype(np.float64), dtype=np.float64, mel=mel64)
            got = logmel_two_stage(w, bits, terms, mel64)
            row['max_abs_logmel_err'][kind] = float(np.abs(got - ref).max())
        rows.append(row)
        print('%-28s %s' % (name, '  '.join('%s %.2e' % kv for kv in row['max_abs_logmel_err'].items())))
    # what the shipped fp32 FFT path achieves on the same signals (oracle fp32 vs fp64)
    base = {}
    for kind in ('white', 'speech', 'lsb', 'dcsine', 'square'):
        w = signals.make(kind, 8000, 7)
        base[kind] = float(np.abs(F.fbank(w.astype(np.float32), mel=mel64.astype(np.float32)).astype(np.float64) -
                                  F.fbank(w.astype(np.float64), dtype=np.float64, mel=mel64)).max())
    print('%-28s %s' % ('fp32 FFT (oracle fp32)', '  '.join('%s %.2e' % kv for kv in base.items())))
    out = {'tolerance': 1e-3, 'signals': '8000 samples each, oracle/signals.py seed 7', 'variants': rows,
           'fp32_fft_reference': base}
    with open(os.path.join(ROOT, 'profiles', 'r02_dft_tc_accuracy.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
