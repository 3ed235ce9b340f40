"""FLAC lists: GPU decode (compressed bytes over PCIe, csrc/oe_flac_gpu.cuh) against host decode (oe_ingest_read's reader
threads), on a batch of the benchmark's shape (256 x U[2,10] s, speech-like synthetic signal, FLAC ratio ~0.53).
Run on the GPU box:  python tools/flac_gpu_bench.py [--batch 256] [--dir /dev/shm/oe_flac]"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openeat_b200 import _lib                                     # noqa: E402
from openeat_b200.ingest import FlacGpuIngest, NativeIngest        # noqa: E402


def speechlike(rng, n, level=1500.0):
    x = rng.normal(0, 1, n + 64)
    for _ in range(3):
        x = np.convolve(x, [0.25, 0.5, 0.25], mode='same')
    x /= x.std()
    env = np.clip(np.sin(2 * np.pi * np.arange(n + 64) / 5000.0 + rng.uniform(0, 6)), 0.02, None)
    return np.clip(np.round(level * x * env)[:n], -32768, 32767).astype(np.int16)


def encode(lib, pcm, block=4096, porder=3):
    out = np.empty(2 * pcm.size + 8192, dtype=np.uint8)
    nbytes = ctypes.c_int64()
    _lib.check(lib.oe_flac_encode(pcm.ctypes.data, pcm.size, 16000, block, porder, out.ctypes.data, out.size, ctypes.byref(nbytes)))
    return out[:nbytes.value].tobytes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--dir', default='/dev/shm/oe_flac')
    ap.add_argument('--reps', type=int, default=20)
    a = ap.parse_args()
    lib = _lib.load()
    os.makedirs(a.dir, exist_ok=True)
    rng = np.random.default_rng(0)
    names, pcm_bytes, flac_bytes, audio_s = [], 0, 0, 0.0
    t0 = time.time()
    for i in range(a.batch):
        n = int(rng.uniform(2.0, 10.0) * 16000)
        x = speechlike(rng, n)
        data = encode(lib, x)
        p = os.path.join(a.dir, 'u%04d.flac' % i)
        with open(p, 'wb') as f:
            f.write(data)
        names.append(p)
        pcm_bytes += 2 * n
        flac_bytes += len(data)
        audio_s += n / 16000.0
    print('corpus: %d files, %.1f audio-s, PCM %.1f MB, FLAC %.1f MB (ratio %.3f), written in %.1f s' % (
        a.batch, audio_s, pcm_bytes / 1e6, flac_bytes / 1e6, flac_bytes / pcm_bytes, time.time() - t0))
    dev = torch.device('cuda:0')
    g = FlacGpuIngest(ring=3)
    c = NativeIngest(ring=3)
    # host side
    for _ in range(3):
        b = g.pack(names)
    t0 = time.time()
    for _ in range(a.reps):
        b = g.pack(names)
    t_pack = (time.time() - t0) / a.reps
    for _ in range(2):
        c.load(names)
    t0 = time.time()
    for _ in range(a.reps):
        buf, offs, lens, rates, loaded, slot = c.load(names)
    t_cpu = (time.time() - t0) / a.reps
    print('host: oe_flac_pack (read + frame index) %.2f ms; oe_ingest_read with host FLAC decode %.2f ms (%d threads) -> %.0f audio-s/s' % (
        t_pack * 1e3, t_cpu * 1e3, c.lib and (os.cpu_count() or 0), audio_s / t_cpu))
    # device side: H2D + decode, CUDA events
    ref = buf.numpy().copy()
    for verify in (True, False):
        for _ in range(3):
            pcm = b.to_device(dev, verify_crc=verify)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        d_comp = b.comp[:b.comp_bytes + 16].to(dev)
        d_frames = b.frames[:b.n_frames * 48].to(dev)
        d_err = torch.zeros(len(names), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev)
        ts = []
        for _ in range(a.reps):
            e[0].record()
            _lib.check(lib.oe_flac_decode_batch(ctypes.c_void_p(d_comp.data_ptr()), b.comp_bytes, ctypes.c_void_p(d_frames.data_ptr()),
                                                b.n_frames, ctypes.c_void_p(pcm.data_ptr()), ctypes.c_void_p(d_err.data_ptr()),
                                                1 if verify else 0, ctypes.c_void_p(st.cuda_stream)))
            e[1].record()
            torch.cuda.synchronize()
            ts.append(e[0].elapsed_time(e[1]))
        assert int(d_err.abs().sum()) == 0
        got = pcm.cpu().numpy()
        for i in range(len(names)):
            assert np.array_equal(got[b.offsets[i]:b.offsets[i] + b.lens[i]], ref[offs[i]:offs[i] + lens[i]])
        print('GPU decode kernel (%d frames, crc %s): median %.3f ms, min %.3f ms -> %.2f M audio-s/s, %.1f GB/s of PCM written' % (
            b.n_frames, verify, float(np.median(ts)), min(ts), audio_s / np.median(ts) / 1e3, pcm_bytes / np.median(ts) / 1e6))
    # H2D alone
    for what, t in (('FLAC', b.comp[:b.comp_bytes + 16]), ('PCM', buf)):
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t.to(dev, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print('H2D of the %s batch: %.1f MB in %.3f ms (%.1f GB/s)' % (what, t.numel() * t.element_size() / 1e6, np.median(ts),
                                                                       t.numel() * t.element_size() / np.median(ts) / 1e6))


if __name__ == '__main__':
    main()
