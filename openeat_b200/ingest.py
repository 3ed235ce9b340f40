"""Native PCM ingest: the decode + pack step in front of the kernels (``oe_ingest_*`` of the C ABI).

The reference decodes one utterance at a time with ``torchaudio.load`` inside DataLoader worker processes
(``openeat/dataset/dataset.py:55-75``, ``openeat/bin/train.py:110-116``).  Here one native call parses the RIFF/WAVE
headers of the whole batch, the utterances are laid out at 8-sample-aligned offsets of ONE pinned buffer, and a pool
of reader threads ``pread``s every utterance straight into its place -- no per-utterance Python, no numpy pack, no GIL
-- so the H2D copy can start from that buffer as it is.
"""
import ctypes
import logging
import os

import numpy as np
import torch

from . import _lib
from ._lib import c_f64p, c_i32p, c_i64p, check
from .frontend import ALIGN, aligned_offsets


def split_entry(entry):
    """'path' or 'path,start,end' (dataset.py:56-58) -> (path, start, end) with start = -1 for a whole file."""
    if ',' not in entry:
        return entry.strip(), -1.0, 0.0
    value = entry.strip().split(',')
    assert len(value) == 1 or len(value) == 3
    if len(value) == 3:
        return value[0], float(value[1]), float(value[2])
    return value[0], -1.0, 0.0


class NativeIngest(object):
    """``load(entries)`` -> (pinned int16 tensor, offsets, lens, sample_rates, loaded, slot).

    The buffer is a slot of a small ring of pinned tensors (grown on demand); call ``release_after(slot, event)`` with
    an event recorded behind the H2D copy that reads it, and the slot is reused only once that copy has completed.
    Entries that cannot be ingested are reported the way the reference reports them (``print`` of the reason and
    ``logging.warning('read utterance ... error')``, dataset.py:108-111) and come back with ``loaded[i] == False``;
    the reason is explicit for formats libsox would have read (FLAC, 24-bit, float)."""

    def __init__(self, threads=0, ring=4):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        threads = int(os.environ.get('OE_INGEST_THREADS', threads))       # 0: one reader per hardware thread
        check(self.lib.oe_ingest_create(int(threads), ctypes.byref(h)))
        self.handle = h
        self._ring = [None] * max(2, int(ring))
        self._events = [None] * len(self._ring)
        self._next = 0

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            self.lib.oe_ingest_destroy(h)
            self.handle = None

    def _slot(self, samples):
        i = self._next
        self._next = (i + 1) % len(self._ring)
        if self._events[i] is not None:
            self._events[i].synchronize()
            self._events[i] = None
        buf = self._ring[i]
        if buf is None or buf.numel() < samples:
            buf = torch.empty(int(samples * 1.25) + 4096, dtype=torch.int16)
            if torch.cuda.is_available():
                buf = buf.pin_memory()
            self._ring[i] = buf
        return i, buf

    def pad_rows(self, src, frames, tmax, dst, pad_row=None):
        """``oe_host_pad_rows``: ragged host rows ``src`` (sum(frames), F) -> zero-padded host tensor ``dst`` (B, tmax, F)
        on this handle's threads (non-temporal stores); ``pad_row`` (F,) float32 replaces the zeros (GlobalCMVN on the
        padding).  Blocking, GIL released; the handle must not have ingest jobs pending."""
        frames = np.ascontiguousarray(frames, dtype=np.int32)
        pr = None if pad_row is None else np.ascontiguousarray(pad_row, dtype=np.float32)
        check(self.lib.oe_host_pad_rows(self.handle, ctypes.c_void_p(src.data_ptr()), frames.ctypes.data_as(c_i32p),
                                        len(frames), int(tmax), int(dst.shape[-1]),
                                        None if pr is None else ctypes.c_void_p(pr.ctypes.data), ctypes.c_void_p(dst.data_ptr())))

    def release_after(self, slot, event):
        self._events[slot] = event

    def submit(self, entries, capacity=None):
        """Starts reading a batch on the handle's driver thread; returns a ticket for ``wait``."""
        n = len(entries)
        parts = [split_entry(e) for e in entries]
        paths = (ctypes.c_char_p * n)(*[p[0].encode() for p in parts])
        starts = np.array([p[1] for p in parts], dtype=np.float64)
        ends = np.array([p[2] for p in parts], dtype=np.float64)
        want = int(capacity or getattr(self, '_seen', 0) or (1 << 22))
        slot, buf = self._slot(want)
        job = ctypes.c_void_p()
        check(self.lib.oe_ingest_submit(self.handle, n, paths, starts.ctypes.data_as(c_f64p), ends.ctypes.data_as(c_f64p),
                                        ctypes.c_void_p(buf.data_ptr()), buf.numel(), ctypes.byref(job)))
        return {'job': job, 'slot': slot, 'buf': buf, 'n': n, 'entries': entries, 'parts': parts}

    def wait(self, ticket, keys=None):
        """Blocks (GIL released) until the batch is in its buffer; same return value as ``load``."""
        n = ticket['n']
        o, l, r, st = c_i64p(), c_i32p(), c_i32p(), c_i32p()
        total = ctypes.c_int64()
        check(self.lib.oe_ingest_wait(ticket['job'], ctypes.byref(o), ctypes.byref(l), ctypes.byref(r), ctypes.byref(st),
                                      ctypes.byref(total)))
        if n:
            offs = np.ctypeslib.as_array(o, shape=(n,)).copy()
            lens = np.ctypeslib.as_array(l, shape=(n,)).copy()
            rates = np.ctypeslib.as_array(r, shape=(n,)).copy()
            status = np.ctypeslib.as_array(st, shape=(n,)).copy()
        else:
            offs, lens, rates, status = np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32)
        self._seen = max(getattr(self, '_seen', 0), int(total.value * 1.25) + 4096)
        if total.value > ticket['buf'].numel():                          # the ring slot was too small: once more, bigger (through the
            self.lib.oe_ingest_job_release(ticket['job'])                # same queue: the driver thread owns the handle's state)
            self._ring[ticket['slot']] = None
            nxt, self._next = self._next, ticket['slot']                 # re-use this very slot; later slots belong to queued jobs
            again = self.submit(ticket['entries'], capacity=int(total.value * 1.25) + 4096)
            self._next = nxt
            return self.wait(again, keys)
        loaded = status == 0
        for i in np.nonzero(~loaded)[0]:                                 # dataset.py:108-111: print, warn, drop
            print(self.lib.oe_ingest_job_error(ticket['job'], int(i)).decode())
            logging.warning('read utterance {} error'.format(keys[i] if keys is not None else ticket['parts'][i][0]))
            lens[i] = 0
        rates[~loaded] = 16000
        self.lib.oe_ingest_job_release(ticket['job'])
        return ticket['buf'][:max(int(total.value), ALIGN)], offs, lens, rates, loaded, ticket['slot']

    def load(self, entries, keys=None, report=True):
        n = len(entries)
        parts = [split_entry(e) for e in entries]
        paths = (ctypes.c_char_p * n)(*[p[0].encode() for p in parts])
        starts = np.array([p[1] for p in parts], dtype=np.float64)
        ends = np.array([p[2] for p in parts], dtype=np.float64)
        lens = np.zeros(n, dtype=np.int32)
        rates = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        sp, ep = starts.ctypes.data_as(c_f64p), ends.ctypes.data_as(c_f64p)
        check(self.lib.oe_ingest_probe(self.handle, n, paths, sp, ep, lens.ctypes.data_as(c_i32p),
                                       rates.ctypes.data_as(c_i32p), status.ctypes.data_as(c_i32p)))
        offs, total = aligned_offsets(lens)
        slot, buf = self._slot(max(total, ALIGN))
        check(self.lib.oe_ingest_read(self.handle, n, paths, sp, ep, ctypes.c_void_p(buf.data_ptr()),
                                      offs.ctypes.data_as(c_i64p), lens.ctypes.data_as(c_i32p),
                                      status.ctypes.data_as(c_i32p)))
        loaded = status == 0
        self.last_errors = {int(i): self.lib.oe_ingest_error(self.handle, int(i)).decode() for i in np.nonzero(~loaded)[0]}
        for i in np.nonzero(~loaded)[0]:                              # dataset.py:108-111: print, warn, drop
            if report:
                print(self.last_errors[int(i)])
                logging.warning('read utterance {} error'.format(keys[i] if keys is not None else parts[i][0]))
            lens[i] = 0
        rates[~loaded] = 16000
        return buf[:max(total, ALIGN)], offs, lens, rates, loaded, slot

    def report_errors(self, keys):
        """Prints what the last ``load(..., report=False)`` held back (dataset.py:108-111 convention)."""
        for i, msg in sorted(getattr(self, 'last_errors', {}).items()):
            print(msg)
            logging.warning('read utterance {} error'.format(keys[i]))


def ingest_batches(item_batches, ingest=None, depth=2):
    """Generator over pre-built batches of ``(key, 'path[,start,end]', tokenid, speed)`` items (what ``AudioDataset``
    yields): up to ``depth`` batches are being read ahead by the handle's native driver thread (``oe_ingest_submit``; no
    Python thread, so nothing competes for the GIL), i.e. file reading overlaps the H2D copy and the kernels of earlier
    batches.  Yields the tuples ``PrefetchingCollator`` takes: ``(pinned_wav, offsets, lens, keys, labels, speeds,
    sample_rates, loaded, release)``; ``release(event)`` hands the ring slot back once ``event`` (recorded behind the H2D
    copy) has completed."""
    ing = ingest or NativeIngest(ring=depth + 3)
    it = iter(item_batches)
    pending = []

    def submit():
        try:
            items = next(it)
        except StopIteration:
            return False
        if len(items) == 1 and isinstance(items[0], list):              # DataLoader-style [batch] wrapping, dataset.py:186-187
            items = items[0]
        pending.append((items, ing.submit([x[1] for x in items])))
        return True

    while len(pending) < depth and submit():
        pass
    while pending:
        items, ticket = pending.pop(0)
        keys = [x[0] for x in items]
        buf, offs, lens, rates, loaded, slot = ing.wait(ticket, keys)
        submit()
        yield (buf, offs, lens, keys, [x[2] for x in items], [x[3] for x in items], rates, loaded,
               (lambda ev, s=slot: ing.release_after(s, ev)))


_default = {}


def default_ingest():
    if 'g' not in _default:
        _default['g'] = NativeIngest()
    return _default['g']
