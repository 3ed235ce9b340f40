"""Accuracy of the hybrid tensor-core / CUDA-core spectrum of the third-generation fbank kernel,
emulated in numpy before any CUDA was written.

Formulation: the 512-point DFT of a frame is factored n = 16 n1 + n2, k = k1 + 32 k2.  Stage 1 (32-point real-input DFTs
over n1, one per residue class n2) runs on the tensor cores as a GEMM whose LEFT operand is the raw int16 PCM, split
EXACTLY into two fp16 pieces (hi = multiple of 16, |hi / 16| <= 2048; lo in [-8, 8]), and whose RIGHT operand is a
constant matrix that folds everything linear in front of the DFT: DC removal (kaldi.py:183-186), pre-emphasis
(193-198), Povey window (201-204), the stage-1 DFT and the inter-stage twiddle W512^(n2 k1); the matrix is split into
two fp16 pieces (22+ significant bits).  Products hi*G1 + hi*G2 + lo*G1, fp32 accumulation.  The frame sum S (DC term)
enters as three extra K slots (exact fp16 pieces of the integer sum).  Stage 2 (sixteen-point complex DFTs over n2) and
power / mel / log run in fp32 on the CUDA cores.

    python tools/ubench/dft_hybrid_accuracy.py       # prints max |log-mel error| vs the fp64 oracle per signal class

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

N1, N2 = 32, 16
WIN, SHIFT = 400, 160


def stage1_matrix(preemph=0.97):
    """G[n, n2, c] (fp64): contribution of frame sample n to column c of class n2, and V[n2, c]: the DC column
    (multiplies the frame sum S).  Columns: 0 = Re Y'[0], 1 = Y[16] (real, untwiddled), 2 k1 / 2 k1 + 1 = Re / Im
    Y'[k1], k1 = 1..15."""
    w = F.povey_window(WIN, np.float64)
    G = np.zeros((WIN, N2, 32))
    V = np.zeros((N2, 32))
    for n2 in range(N2):
        for n1 in range((WIN - n2 + 15) // 16):
            n = 16 * n1 + n2
            for k1 in range(17):
                t = w[n] * np.exp(-2j * np.pi * (n1 * k1 / 32.0))
                if k1 != 16:
                    t = t * np.exp(-2j * np.pi * n2 * k1 / 512.0)
                cols = [(0, t.real)] if k1 == 0 else [(1, t.real)] if k1 == 16 else [(2 * k1, t.real), (2 * k1 + 1, t.imag)]
                for c, v in cols:
                    # z[n] = w[n] ((x[n] - mu) - p (x[n-1] - mu)),  n = 0: x[-1] := x[0]
                    G[n, n2, c] += v
                    G[max(n - 1, 0), n2, c] -= preemph * v
                    V[n2, c] -= (1.0 - preemph) * v / WIN
    return G, V


def split_f16(m, scale2=2048.0):
    """two fp16 pieces: m ~= p1 + p2 / scale2 (the residual is stored scaled up so it stays a normal fp16)."""
    p1 = m.astype(np.float16)
    p2 = ((m - p1.astype(np.float64)) * scale2).astype(np.float16)
    return p1.astype(np.float64), p2.astype(np.float64) / scale2


def logmel_hybrid(wave, mel64, center=False, drop_lo_g2=True):
    x = F.frames_of(wave.astype(np.float64))                       # (m, 400) integer-valued
    m = x.shape[0]
    if center:                                                     # subtract an integer constant per frame group (exact)
        c = np.round(x.mean(axis=1, keepdims=True))
        x = x - c
    hi = x.astype(np.float16).astype(np.float64)                  # floating split: 11 significant bits, exact remainder
    lo = x - hi
    assert np.abs(lo).max() <= 32 and np.all(lo == lo.astype(np.float16))
    S = x.sum(axis=1)                                              # exact integer, |S| < 2^24
    s1 = np.round(S / 4096.0) * 4096.0
    s2 = np.round((S - s1) / 2.0) * 2.0
    s3 = S - s1 - s2
    G, V = stage1_matrix()
    G1, G2 = split_f16(G.reshape(WIN, -1))
    V1, V2 = split_f16(V.reshape(1, -1))
    f32 = np.float32
    acc = (hi.astype(f32) @ G1.astype(f32)).astype(f32)
    acc = acc + (hi.astype(f32) @ G2.astype(f32)).astype(f32)
    acc = acc + (lo.astype(f32) @ G1.astype(f32)).astype(f32)
    if not drop_lo_g2:
        acc = acc + (lo.astype(f32) @ G2.astype(f32)).astype(f32)
    acc = acc + (s1[:, None].astype(f32) * (V1 + V2).astype(f32)) + (s2[:, None].astype(f32) * (V1 + V2).astype(f32)) \
        + (s3[:, None].astype(f32) * V1.astype(f32))
    Y = acc.astype(f32).reshape(m, N2, 32)
    # stage 2 in fp32: per k1 a 16-point complex DFT over n2
    Yc = np.zeros((m, N2, 17), np.complex64)
    Yc[:, :, 0] = Y[:, :, 0]
    tw16 = np.exp(-2j * np.pi * np.arange(N2) * 16 / 512.0).astype(np.complex64)
    Yc[:, :, 16] = Y[:, :, 1] * tw16[None]
    for k1 in range(1, 16):
        Yc[:, :, k1] = Y[:, :, 2 * k1] + 1j * Y[:, :, 2 * k1 + 1]
    D = np.exp(-2j * np.pi * np.outer(np.arange(N2), np.arange(16)) / 16.0).astype(np.complex64)    # [n2, k2]
    X = np.einsum('mnk,nq->mkq', Yc, D).astype(np.complex64)       # [frame, k1, k2] = X[k1 + 32 k2]
    power = np.zeros((m, 257), np.float64)
    for k1 in range(17):
        for k2 in range(16):
            k = k1 + 32 * k2
            kk = k if k <= 256 else 512 - k
            if k1 in (0, 16) and k > 256:
                continue
            power[:, kk] = np.abs(X[:, k1, k2].astype(np.complex128)) ** 2
    melp = np.concatenate([mel64, np.zeros((mel64.shape[0], 1))], axis=1)
    return np.log(np.maximum(power @ melp.T, float(F.EPS_F32)))


def main():
    mel64 = F.mel_banks(80, 512, 16000.0, 20.0, 0.0, np.float32).astype(np.float64)
    kinds = ('white', 'speech', 'lsb', 'dcsine', 'square')
    out = {}
    for name, kw in (('hybrid fp16 2x2 pieces, 3 products', {}), ('+ lo*G2 (4 products)', {'drop_lo_g2': False}),
                     ('3 products, per-frame integer centring', {'center': True})):
        row = {}
        for kind in kinds:
            w = signals.make(kind, 8000, 7)
            ref = F.fbank(w.astype(np.float64), dtype=np.float64, mel=mel64)
            row[kind] = float(np.abs(logmel_hybrid(w, mel64, **kw) - ref).max())
        out[name] = row
        print('%-42s %s' % (name, '  '.join('%s %.2e' % kv for kv in row.items())))
    base = {}
    for kind in kinds:
        w = signals.make(kind, 8000, 7)
        base[kind] = float(np.abs(F.fbank(w.astype(np.float32), mel=mel64.astype(np.float32)).astype(np.float64) -
                                  F.fbank(w.astype(np.float64), dtype=np.float64, mel=mel64)).max())
    out['fp32 FFT (oracle fp32)'] = base
    print('%-42s %s' % ('fp32 FFT (oracle fp32)', '  '.join('%s %.2e' % kv for kv in base.items())))
    with open(os.path.join(ROOT, 'profiles', 'r02_dft_hybrid_accuracy.json'), 'w') as f:
        json.dump({'tolerance': 1e-3, 'signals': '8000 samples each, oracle/signals.py seed 7', 'max_abs_logmel_err': out}, f, indent=1)


if __name__ == '__main__':
    main()
