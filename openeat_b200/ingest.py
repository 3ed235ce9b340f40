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

    def _request(self, entries):
        """entries -> (n, parts, char** array, starts, ends).  The pointer array is built with numpy over ONE encoded blob
        (a ctypes array of 256 c_char_p costs ~90 us of the collate thread per batch); `parts` keeps the blob alive."""
        n = len(entries)
        joined = '\0'.join(entries)
        if ',' in joined or ' ' in joined or '\n' in joined or '\t' in joined:    # segments / stray whitespace: entry by entry
            parts = [split_entry(e) for e in entries]
            starts = np.array([p[1] for p in parts], dtype=np.float64)
            ends = np.array([p[2] for p in parts], dtype=np.float64)
            joined = '\0'.join(p[0] for p in parts)
        else:
            parts = [(e, -1.0, 0.0) for e in entries]
            starts = np.full(n, -1.0, dtype=np.float64)
            ends = np.zeros(n, dtype=np.float64)
        blob = (joined + '\0').encode()
        ptrs = np.zeros(max(n, 1), dtype=np.uint64)
        if n:
            base = ctypes.cast(ctypes.c_char_p(blob), ctypes.c_void_p).value
            ends_at = np.flatnonzero(np.frombuffer(blob, dtype=np.uint8) == 0)
            assert len(ends_at) == n, 'a path contains a NUL byte'
            ptrs[0] = base
            ptrs[1:n] = base + ends_at[:-1].astype(np.uint64) + np.uint64(1)
        paths = ptrs.ctypes.data_as(ctypes.POINTER(ctypes.c_char_p))
        parts.append((blob, ptrs))               # keep-alive for the duration of the native call
        return n, parts, paths, starts, ends

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
        n, parts, paths, starts, ends = self._request(entries)
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
        n, parts, paths, starts, ends = self._request(entries)
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


# ---------------------------------------------------------------------------------------------------------------------
# FLAC lists: compressed bytes over PCIe, decoded by the GPU (csrc/oe_flac_gpu.cuh)
# ---------------------------------------------------------------------------------------------------------------------
FRAME_BYTES = 48                                    # sizeof(oe_flac_frame), include/openeat_frontend.h
FLAC_ERRORS = {1: 'a frame does not end where the next one starts', 2: 'frame CRC-16 mismatch',
               4: 'predictor order above 12 (host decoder needed)', 8: 'reserved code in a subframe',
               16: 'the bit stream runs past the end of the buffer'}


class FlacBatch(object):
    """One batch of FLAC files packed for the GPU decoder (``FlacGpuIngest.pack``): ``comp`` (pinned uint8: the files,
    each 16-byte aligned) and ``frames`` (pinned uint8 holding ``n_frames`` ``oe_flac_frame`` records) cross PCIe instead
    of the PCM; ``to_device`` enqueues the two copies and the decode kernel on the current stream and returns the packed
    int16 PCM tensor ``collate_packed`` takes (utterance i at ``offsets[i]``, ``lens[i]`` samples)."""

    def __init__(self, owner, slot, comp, frames, comp_bytes, n_frames, total, offsets, lens, rates, loaded, keys):
        self.owner, self.slot = owner, slot
        self.comp, self.frames = comp, frames
        self.comp_bytes, self.n_frames, self.total = int(comp_bytes), int(n_frames), int(total)
        self.offsets, self.lens, self.rates, self.loaded, self.keys = offsets, lens, rates, loaded, keys
        self.errors = None                          # pinned int32 per entry, valid once `event` has completed
        self.paths = None
        self.event = None

    is_cuda = False                                 # PrefetchingCollator: not a device tensor yet

    @property
    def h2d_bytes(self):
        return self.comp_bytes + 16 + self.n_frames * FRAME_BYTES

    def to_device(self, device, verify_crc=True, wait=True):
        lib = self.owner.lib
        n = len(self.lens)
        pcm = torch.empty(max(self.total, ALIGN), dtype=torch.int16, device=device)
        if self.n_frames == 0:                   # nothing decodable in the batch: still an event for whoever orders behind it
            self.event = torch.cuda.Event()
            self.event.record(torch.cuda.current_stream(device))
            self.errors = torch.zeros(max(n, 1), dtype=torch.int32)
            return pcm
        d_comp = self.comp[:self.comp_bytes + 16].to(device, non_blocking=True)
        d_frames = self.frames[:self.n_frames * FRAME_BYTES].to(device, non_blocking=True)
        d_err = torch.zeros(max(n, 1), dtype=torch.int32, device=device)
        # copies on the caller's (current) stream, the kernel on the ingest's decode stream behind them: the next batch's copy
        # does not queue behind this batch's decode
        cur = torch.cuda.current_stream(device)
        stream = self.owner.decode_stream(device)
        copied = torch.cuda.Event()
        copied.record(cur)
        stream.wait_event(copied)
        check(lib.oe_flac_decode_batch(ctypes.c_void_p(d_comp.data_ptr()), self.comp_bytes, ctypes.c_void_p(d_frames.data_ptr()),
                                       self.n_frames, ctypes.c_void_p(pcm.data_ptr()), ctypes.c_void_p(d_err.data_ptr()),
                                       1 if verify_crc else 0, ctypes.c_void_p(stream.cuda_stream)))
        self.errors = torch.empty(max(n, 1), dtype=torch.int32).pin_memory()
        with torch.cuda.stream(stream):
            self.errors.copy_(d_err, non_blocking=True)
        for t in (d_comp, d_frames, d_err, pcm):
            t.record_stream(stream)
        self.event = torch.cuda.Event()
        self.event.record(stream)
        if wait:
            cur.wait_event(self.event)           # stream-ordered for the caller: the PCM is ready for whatever it enqueues next
        self.owner._watch(self)
        return pcm

    def drop_failed(self, lens, loaded, keys=None):
        """Waits for the decode; entries whose stream failed the kernel's checks are reported and dropped with the
        reference's convention (print, ``logging.warning('read utterance ... error')``, dataset.py:108-111): their
        ``lens`` become 0 and ``loaded`` False (both updated in place and returned as ``lens``)."""
        if self.event is None:
            return lens
        self.event.synchronize()
        bad = np.nonzero(self.errors.numpy()[:len(self.lens)])[0]
        for i in bad:
            print('%s: FLAC stream rejected by the GPU decoder: %s' % (
                self.paths[i] if self.paths is not None else i, ', '.join(v for k, v in FLAC_ERRORS.items() if int(self.errors[i]) & k)))
            logging.warning('read utterance {} error'.format(keys[i] if keys is not None else i))
            lens[i] = 0
            loaded[i] = False
        self.errors[:len(self.lens)] = 0         # reported: the owner's deferred check stays quiet
        return lens

    def check(self):
        """Waits for the decode and raises ``FrontendError`` naming every entry whose stream failed the GPU decoder's
        checks (end-of-frame position, CRC-16) -- corrupted audio is never passed on silently."""
        if self.event is None:
            return
        self.event.synchronize()
        bad = np.nonzero(self.errors.numpy()[:len(self.lens)])[0]
        if len(bad):
            raise _lib.FrontendError('; '.join('%s: %s' % (self.keys[i] if self.keys is not None else i, ', '.join(
                v for k, v in FLAC_ERRORS.items() if int(self.errors[i]) & k)) for i in bad))


class FlacGpuIngest(NativeIngest):
    """``pack(entries)`` -> ``FlacBatch``: the reader pool ``pread``s the FLAC files of a batch into one pinned buffer and
    indexes their frames (``oe_flac_pack``); nothing is decoded on the host.  Entries the GPU decoder does not take
    (stereo, 24-bit, streamed encodes without a length) are reported and dropped with the reference's convention
    (dataset.py:108-111); decode them with ``NativeIngest`` / ``read_wav`` instead."""

    def __init__(self, threads=0, ring=4):
        super(FlacGpuIngest, self).__init__(threads, ring)
        self._comp = [None] * len(self._ring)
        self._frames = [None] * len(self._ring)
        self._watched = []
        self._decode_streams = {}

    def decode_stream(self, device):
        key = str(device)
        if key not in self._decode_streams:
            self._decode_streams[key] = torch.cuda.Stream(device=device)
        return self._decode_streams[key]

    def _watch(self, batch):
        """Batches whose decode has finished are checked when the next one is launched (no extra synchronisation)."""
        keep = []
        for b in self._watched:
            if b.event.query():
                b.check()
            else:
                keep.append(b)
        keep.append(batch)
        self._watched = keep

    def _buffers(self, comp_bytes, n_frames):
        i = self._next
        self._next = (i + 1) % len(self._ring)
        if self._events[i] is not None:
            self._events[i].synchronize()
            self._events[i] = None

        def grow(buf, need):
            if buf is None or buf.numel() < need:
                buf = torch.empty(int(need * 1.25) + 4096, dtype=torch.uint8)
                if torch.cuda.is_available():
                    buf = buf.pin_memory()
            return buf
        self._comp[i] = grow(self._comp[i], comp_bytes + 16)
        self._frames[i] = grow(self._frames[i], n_frames * FRAME_BYTES)
        return i, self._comp[i], self._frames[i]

    def _finish(self, n, parts, keys, report, slot, comp, frames, cb, nf, total, offs, lens, rates, status, error_of):
        self._seen_flac = (max(getattr(self, '_seen_flac', (0, 0))[0], cb), max(getattr(self, '_seen_flac', (0, 0))[1], nf))
        loaded = status == 0
        for i in np.nonzero(~loaded)[0]:                               # dataset.py:108-111: print, warn, drop
            if report:
                print(error_of(int(i)))
                logging.warning('read utterance {} error'.format(keys[i] if keys is not None else parts[i][0]))
            lens[i] = 0
        rates[~loaded] = 16000
        b = FlacBatch(self, slot, comp, frames, cb, nf, total, offs, lens, rates, loaded, keys)
        b.paths = [p[0] for p in parts[:n]]
        return b

    def pack(self, entries, keys=None, report=True):
        n, parts, paths, starts, ends = self._request(entries)
        coffs, offs = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        lens, rates, status = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        cb, nf, total = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        want = getattr(self, '_seen_flac', (1 << 20, 1 << 12))
        for attempt in range(2):
            slot, comp, frames = self._buffers(*want)
            rc = self.lib.oe_flac_pack(self.handle, n, paths, starts.ctypes.data_as(c_f64p), ends.ctypes.data_as(c_f64p),
                                       ctypes.c_void_p(comp.data_ptr()), comp.numel(), ctypes.c_void_p(frames.data_ptr()),
                                       frames.numel() // FRAME_BYTES, coffs.ctypes.data_as(c_i64p), offs.ctypes.data_as(c_i64p),
                                       lens.ctypes.data_as(c_i32p), rates.ctypes.data_as(c_i32p), status.ctypes.data_as(c_i32p),
                                       ctypes.byref(cb), ctypes.byref(nf), ctypes.byref(total))
            if rc != _lib.OE_ERR_WORKSPACE:
                break
            want = (cb.value, nf.value)                                # too small: once more with what the call asked for
            self._next = slot
        check(rc)
        return self._finish(n, parts, keys, report, slot, comp, frames, cb.value, nf.value, total.value, offs, lens, rates, status,
                            lambda i: self.lib.oe_ingest_error(self.handle, i).decode())

    def submit(self, entries, keys=None):
        """Starts packing a batch on the handle's native driver thread (``oe_flac_submit``: no Python thread, nothing competes
        for the GIL); returns a ticket for ``wait``."""
        n, parts, paths, starts, ends = self._request(entries)
        want = getattr(self, '_seen_flac', (1 << 20, 1 << 12))
        slot, comp, frames = self._buffers(*want)
        job = ctypes.c_void_p()
        check(self.lib.oe_flac_submit(self.handle, n, paths, starts.ctypes.data_as(c_f64p), ends.ctypes.data_as(c_f64p),
                                      ctypes.c_void_p(comp.data_ptr()), comp.numel(), ctypes.c_void_p(frames.data_ptr()),
                                      frames.numel() // FRAME_BYTES, ctypes.byref(job)))
        return {'job': job, 'slot': slot, 'comp': comp, 'frames': frames, 'entries': entries, 'keys': keys, 'n': n, 'parts': parts}

    def wait(self, ticket, report=True):
        """Blocks (GIL released) until the batch is packed; returns the ``FlacBatch``."""
        n = ticket['n']
        o, l, r, st = c_i64p(), c_i32p(), c_i32p(), c_i32p()
        cb, nf, total = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        rc = self.lib.oe_flac_wait(ticket['job'], ctypes.byref(o), ctypes.byref(l), ctypes.byref(r), ctypes.byref(st),
                                   ctypes.byref(cb), ctypes.byref(nf), ctypes.byref(total))
        if rc == _lib.OE_ERR_WORKSPACE:                                  # the ring slot was too small: once more, bigger, through the
            self.lib.oe_ingest_job_release(ticket['job'])                # same queue (the driver thread owns the handle's state)
            seen = getattr(self, '_seen_flac', (0, 0))
            self._seen_flac = (max(seen[0], int(cb.value * 1.25) + 65536), max(seen[1], int(nf.value * 1.25) + 256))
            nxt, self._next = self._next, ticket['slot']                 # re-use this very slot; later slots belong to queued jobs
            again = self.submit(ticket['entries'], ticket['keys'])
            self._next = nxt
            return self.wait(again, report)
        if rc != 0:
            self.lib.oe_ingest_job_release(ticket['job'])
            check(rc)
        take = lambda p, dt: np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)   # noqa: E731
        offs, lens, rates, status = take(o, np.int64), take(l, np.int32), take(r, np.int32), take(st, np.int32)
        job = ticket['job']
        b = self._finish(n, ticket['parts'], ticket['keys'], report, ticket['slot'], ticket['comp'], ticket['frames'], cb.value,
                         nf.value, total.value, offs, lens, rates, status, lambda i: self.lib.oe_ingest_job_error(job, i).decode())
        self.lib.oe_ingest_job_release(job)
        return b


def flac_gpu_batches(item_batches, ingest=None, depth=3, workers=2, threads=0):
    """``ingest_batches`` for FLAC lists decoded on the GPU: yields the tuples ``PrefetchingCollator`` takes with a
    ``FlacBatch`` in the place of the pinned PCM tensor (the collator calls its ``to_device`` on the copy stream).  The next
    ``depth`` batches are being packed ahead on the native driver threads of ``workers`` ``FlacGpuIngest`` handles
    (``oe_flac_submit`` / ``oe_flac_wait``; consecutive batches alternate between the handles, each handle packs one batch at
    a time; ``threads`` reader threads in total, 0 = one per core).  No Python threads: helper threads that need the GIL
    between their native calls were measured to stall the collate thread for whole switch intervals."""
    if ingest is not None:
        ings = [ingest]
    else:
        total = int(threads) or len(os.sched_getaffinity(0))
        workers = max(1, min(int(workers), total))
        ings = [FlacGpuIngest(threads=max(1, total // workers), ring=depth + 4) for _ in range(workers)]
    it = iter(item_batches)
    pending = []
    count = [0]

    def submit():
        try:
            items = next(it)
        except StopIteration:
            return False
        if len(items) == 1 and isinstance(items[0], list):
            items = items[0]
        ing = ings[count[0] % len(ings)]
        count[0] += 1
        pending.append((items, ing, ing.submit([x[1] for x in items], [x[0] for x in items])))
        return True

    while len(pending) < depth and submit():
        pass
    try:
        while pending:
            items, ing, ticket = pending.pop(0)
            b = ing.wait(ticket)
            submit()
            yield (b, b.offsets, b.lens, [x[0] for x in items], [x[2] for x in items], [x[3] for x in items], b.rates, b.loaded,
                   (lambda ev, s=b.slot, g=ing: g.release_after(s, ev)))
    finally:
        for _, ing, ticket in pending:           # packs in flight finish before their handles (and pinned buffers) can go
            try:
                ing.lib.oe_flac_wait(ticket['job'], None, None, None, None, None, None, None)
                ing.lib.oe_ingest_job_release(ticket['job'])
            except Exception:                    # interpreter shutdown: modules are already gone
                pass


def default_flac_ingest():
    if 'flac' not in _default:
        _default['flac'] = FlacGpuIngest()
    return _default['flac']
