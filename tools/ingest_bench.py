"""Developer timing of the native ingest alone (host only): 256 wav files of 2-10 s in tmpfs, probe + read, by thread count."""
import ctypes
import os
import shutil
import sys
import tempfile
import time
import wave

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openeat_b200._lib import c_f64p, c_i32p, c_i64p   # noqa: E402
from openeat_b200.frontend import aligned_offsets      # noqa: E402
from openeat_b200.ingest import NativeIngest           # noqa: E402

d = tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
rng = np.random.default_rng(0)
lens0 = np.round(rng.uniform(2, 10, 256) * 16000).astype(int)
entries = []
for i, n in enumerate(lens0):
    p = '%s/u%d.wav' % (d, i)
    with wave.open(p, 'wb') as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(rng.integers(-3000, 3000, n).astype('<i2').tobytes())
    entries.append(p)
n = len(entries)
try:
    import torch
    pin = torch.cuda.is_available()
except Exception:
    pin = False
print('cpus', len(os.sched_getaffinity(0)), 'pinned destination', pin)
for th in (1, 2, 4, 8, 16, 32):
    ing = NativeIngest(threads=th)
    paths = (ctypes.c_char_p * n)(*[e.encode() for e in entries])
    starts, ends = np.full(n, -1.0), np.zeros(n)
    lens, rates, status = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    if pin:
        t = torch.empty(int(lens0.sum() + 8 * n), dtype=torch.int16).pin_memory()
        ptr = t.data_ptr()
    else:
        t = np.zeros(int(lens0.sum() + 8 * n), np.int16)
        ptr = t.ctypes.data

    def once():
        t0 = time.perf_counter()
        ing.lib.oe_ingest_probe(ing.handle, n, paths, starts.ctypes.data_as(c_f64p), ends.ctypes.data_as(c_f64p),
                                lens.ctypes.data_as(c_i32p), rates.ctypes.data_as(c_i32p), status.ctypes.data_as(c_i32p))
        t1 = time.perf_counter()
        offs, total = aligned_offsets(lens)
        ing.lib.oe_ingest_read(ing.handle, n, paths, starts.ctypes.data_as(c_f64p), ends.ctypes.data_as(c_f64p),
                               ctypes.c_void_p(ptr), offs.ctypes.data_as(c_i64p), lens.ctypes.data_as(c_i32p),
                               status.ctypes.data_as(c_i32p))
        return t1 - t0, time.perf_counter() - t1
    once()
    r = [once() for _ in range(10)]
    pr, rd = np.median([a for a, b in r]), np.median([b for a, b in r])
    print('%2d threads: probe %.2f ms  read %.2f ms  (%.1f GB/s)' % (th, 1e3 * pr, 1e3 * rd, 2 * lens0.sum() / rd / 1e9))
shutil.rmtree(d)
