"""Drop-in for the front-end part of ``openeat/dataset/dataset.py``: ``_extract_feature`` and
``audio_collate_func`` with the reference's signatures, output structure, sort order, Python-``random``
call order and drop-on-error convention -- the numeric work is ONE fused launch
sequence per batch on the GPU instead of a per-utterance CPU loop.

Differences a caller can observe (all deliberate):
  * tensors come back on the GPU by default (``output_device='cpu'`` restores the reference's CPU
    tensors); ``features`` are the same zero-padded (B, Tmax, F) fp32 tensor either way;
  * audio decoding is built in: RIFF/WAVE files (integer PCM 8/16/24/32 bit, IEEE float) and FLAC streams (RFC 9639,
    the LibriSpeech corpus; decoded natively by ``oe_flac_decode``) -- the reference needs libsox;
    an item's second field may also be an in-memory int16 / fp32 array or ``(array, sample_rate)``;
  * ``data_type != 'wav'`` reads binary Kaldi archives with the built-in ``openeat_b200.kaldi_io.read_mat``
    (the reference imports the third-party ``kaldi_io``); ``feature_dither`` draws its amplitude with the
    reference's ``random.uniform`` call and the per-cell uniforms with Philox on the GPU (no value parity);
  * speed perturb uses the torchaudio sinc resampler semantics (libsox is not reproducible here);
  * optional extras the reference applies later on the device can be fused in:
    ``global_cmvn=(mean, istd)`` (GlobalCMVN, encoder.py:221-222) and ``cmvn_stats`` accumulation.
"""
import collections
import logging
import os
import random
import struct

import numpy as np
import torch

from . import kaldi_io, planner
from .frontend import default_frontend, pack_waveforms, speed_ratio

IGNORE_ID = -1  # openeat/utils/common.py:24


def _riff_chunks(f):
    """(format tag, channels, sample rate, bits, data offset, data bytes) of a RIFF/WAVE file object."""
    head = f.read(12)
    if len(head) < 12 or head[:4] != b'RIFF' or head[8:12] != b'WAVE':
        raise ValueError('not a RIFF/WAVE or FLAC file')
    size = f.seek(0, 2)
    fmt = None
    pos = 12
    while pos + 8 <= size:
        f.seek(pos)
        cid, csz = struct.unpack('<4sI', f.read(8))
        if cid == b'fmt ':
            body = f.read(min(csz, 40))
            tag, nch, sr, _, _, bits = struct.unpack('<HHIIHH', body[:16])
            if tag == 0xFFFE and len(body) >= 26:                  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID
                tag = struct.unpack('<H', body[24:26])[0]
            fmt = (tag, nch, sr, bits)
        elif cid == b'data':
            if fmt is None:
                raise ValueError('data chunk before fmt chunk')
            avail = size - (pos + 8)
            nbytes = avail if csz in (0, 0xFFFFFFFF) else min(csz, avail)   # streamed files write 0 / 0xFFFFFFFF sizes
            return fmt + (pos + 8, nbytes)
        pos += 8 + csz + (csz & 1)
    raise ValueError('no data chunk')


def read_wav(path, start=None, end=None):
    """dataset.py:62-75 (``torchaudio.load`` then ``* (1 << 15)``) for RIFF/WAVE files and FLAC streams: returns (samples of channel 0
    on the int16 scale, sample_rate).  16-bit PCM comes back as int16 (the values the reference holds as fp32); 8-bit
    (unsigned), 24- and 32-bit PCM and IEEE float come back as float32 ``normalised * 32768`` with torchaudio's
    normalisation (``(s - 128) / 2^7``, ``s / 2^23``, ``s / 2^31``, float as is).  ``start`` / ``end`` are seconds
    (segmented wav.scp entries ``path,start,end``: ``frame_offset = int(start * sr)``, ``num_frames = int(end * sr) -
    frame_offset``)."""
    with open(path, 'rb') as f:
        return decode_wav(f, start, end, path)


def _decode_flac(data, start, end, name):
    """A whole FLAC stream in memory -> (channel 0 on the int16 scale, sample_rate); the decoder is the library's
    ``oe_flac_decode`` (csrc/oe_flac.h: frame CRCs always checked).  16-bit streams come back as int16, other sample
    sizes as float32 ``s / 2^(bits-1) * 32768`` (torchaudio's normalisation followed by dataset.py:75)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    sr, nch, bits, total = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
    if lib.oe_flac_info(buf.ctypes.data, buf.size, ctypes.byref(sr), ctypes.byref(nch), ctypes.byref(bits), ctypes.byref(total)):
        raise ValueError('%s: %s' % (name, lib.oe_last_error().decode()))
    sr, bits, total = sr.value, bits.value, total.value
    first, count = 0, total
    if start is not None:
        s = int(float(start) * sr)
        e = int(float(end) * sr)
        first = min(max(s, 0), total)
        count = max(0, min(e - s, total - first))
    out = np.empty(max(count, 1), dtype=np.int32)
    if lib.oe_flac_decode(buf.ctypes.data, buf.size, 0, first, count, out.ctypes.data, 0, None):
        raise ValueError('%s: %s' % (name, lib.oe_last_error().decode()))
    out = out[:count]
    if bits == 16:
        return out.astype(np.int16), sr
    return out.astype(np.float32) / np.float32(2.0 ** (bits - 1)) * np.float32(1 << 15), sr


def _is_flac(f):
    """True when the (seekable) file object holds a native FLAC stream, possibly behind an ID3v2 tag; rewinds."""
    head = f.read(10)
    ok = head[:4] == b'fLaC'
    if not ok and len(head) == 10 and head[:3] == b'ID3' and not any(b & 0x80 for b in head[6:10]):
        f.seek(10 + (head[6] << 21 | head[7] << 14 | head[8] << 7 | head[9]) + (10 if head[5] & 0x10 else 0))
        ok = f.read(4) == b'fLaC'
    f.seek(0)
    return ok


def decode_wav(f, start=None, end=None, name='<wav>'):
    """``read_wav`` on an open (seekable) binary file object, e.g. a member of a shard tar."""
    if _is_flac(f):
        return _decode_flac(f.read(), start, end, name)
    try:
        tag, nch, sr, bits, off, nbytes = _riff_chunks(f)
    except (ValueError, struct.error) as e:
        raise ValueError('%s: %s' % (name, e))
    if tag not in (1, 3) or (tag == 1 and bits not in (8, 16, 24, 32)) or (tag == 3 and bits not in (32, 64)) or nch < 1:
        raise ValueError('%s: wav format tag %d with %d-bit samples is not supported (integer PCM 8/16/24/32 bit, '
                         'IEEE float 32/64 bit)' % (name, tag, bits))
    width = bits // 8 * nch
    total = nbytes // width
    first, count = 0, total
    if start is not None:
        s = int(float(start) * sr)
        e = int(float(end) * sr)
        first = min(max(s, 0), total)
        count = max(0, min(e - s, total - first))
    f.seek(off + first * width)
    raw = f.read(count * width)
    if tag == 1 and bits == 16:
        pcm = np.frombuffer(raw, dtype='<i2')
        return (pcm.reshape(-1, nch)[:, 0] if nch > 1 else pcm), sr
    if tag == 3:
        x = np.frombuffer(raw, dtype='<f4' if bits == 32 else '<f8').reshape(-1, nch)[:, 0].astype(np.float32)
    elif bits == 8:
        x = (np.frombuffer(raw, dtype=np.uint8).reshape(-1, nch)[:, 0].astype(np.float32) - 128.0) / 128.0
    elif bits == 24:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, nch, 3)[:, 0, :].astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float32) / np.float32(1 << 23)
    else:
        x = np.frombuffer(raw, dtype='<i4').reshape(-1, nch)[:, 0].astype(np.float32) / np.float32(2.0 ** 31)
    return x * np.float32(1 << 15), sr


def _decodable(entry):
    """True when ``read_wav`` can decode the file of a ``path[,start,end]`` entry."""
    try:
        with open(entry.strip().split(',')[0], 'rb') as f:
            if _is_flac(f):
                return True
            tag, nch, _, bits, _, _ = _riff_chunks(f)
        return nch >= 1 and ((tag == 1 and bits in (8, 16, 24, 32)) or (tag == 3 and bits in (32, 64)))
    except (OSError, ValueError, struct.error):
        return False


def _load_item(x):
    """(samples, sample_rate) for one batch item ``(key, path_or_array, tokenid, speed)``."""
    wav = x[1]
    if isinstance(wav, str):
        value = wav.strip().split(",")
        # 1 for general wav.scp, 3 for segmented wav.scp (dataset.py:56-58)
        assert len(value) == 1 or len(value) == 3
        if len(value) == 3:
            return read_wav(value[0], value[1], value[2])
        return read_wav(value[0])
    if isinstance(wav, tuple):
        return np.asarray(wav[0]), int(wav[1])
    return np.asarray(wav), 16000


class _Plan(object):
    """Host-side description of one batch after all random draws, before any GPU work.  ``src[i]`` is the
    position, in the packed input buffer, of the utterance that ends up i-th after the length sort;
    ``stage1`` / ``stage2`` are the (orig, new) resampling ratios of the resample_rate and speed stages
    (0, 0 = none)."""
    __slots__ = ('keys', 'labels', 'src', 'stage1', 'stage2', 'frames', 'sample_rate', 'wav_dither', 'resampler')


def _ceil_ratio(n, orig, new):
    """torchaudio functional.py:1427: ceil(new * n / orig), vectorised."""
    return (new * n + orig - 1) // orig


def _plan_batch(keys_in, labels_in, nsamples, sample_rates, speeds_in, conf, target_rate=16000, loaded=None):
    """The host half of _extract_feature (dataset.py:47-118) for utterances that are already in memory:
    decide speeds with the reference's RNG call order (natively, openeat_b200.planner), compute frame
    counts, drop failures (printing the reference's message), sort by length descending.  ``loaded[i]``
    False marks an utterance whose load failed before the reference would have drawn any random number."""
    B = len(keys_in)
    n = np.asarray(nsamples, dtype=np.int64).reshape(B)
    sr = np.asarray(sample_rates, dtype=np.int64).reshape(B)
    active = np.ones(B, dtype=bool) if loaded is None else np.asarray(loaded, dtype=bool)
    fe = default_frontend(conf['mel_bins'], target_rate)
    speed = planner.plan_speeds(conf.get('speed_perturb_rate', 0.5), conf.get('speeds', None),
                                np.asarray(speeds_in, dtype=np.float64).reshape(B), active)   # dataset.py:87-89
    stage1 = np.zeros((B, 2), dtype=np.int64)
    stage2 = np.zeros((B, 2), dtype=np.int64)
    rr = sr.copy() if 'resample_rate' not in conf else np.full(B, int(conf['resample_rate']), dtype=np.int64)
    for s in np.unique(sr[active & (rr != sr)]):                  # dataset.py:77-84
        sel = active & (sr == s) & (rr != sr)
        r = int(rr[sel][0])
        g = int(np.gcd(int(s), r))
        stage1[sel] = (int(s) // g, r // g)
        n[sel] = _ceil_ratio(n[sel], int(s) // g, r // g)
    for v in np.unique(speed[active & (speed != 1.0)]):           # dataset.py:90-91
        sel = active & (speed == v)
        for r in np.unique(rr[sel]):
            sel2 = sel & (rr == r)
            ratio = speed_ratio(float(v), int(r))
            stage2[sel2] = ratio
            n[sel2] = _ceil_ratio(n[sel2], ratio[0], ratio[1])
    frames = fe.num_frames_array(n)
    bad = active & ((rr != target_rate) | (frames == 0))
    for i in np.nonzero(bad)[0]:                                  # dataset.py:108-111: print, warn, drop
        if rr[i] != target_rate:
            print('sample rate %d is not supported by this front-end build (needs %d; set resample_rate)' % (rr[i], target_rate))
        else:   # kaldi.py:142 asserts 2 <= window_size <= len(waveform); the reference prints it and drops
            print('choose a window size 400 that is [2, %d]' % n[i])
        logging.warning('read utterance {} error'.format(keys_in[i]))
    keep = np.nonzero(active & ~bad)[0]
    order = np.argsort(frames[keep])[::-1] if len(keep) else np.zeros(0, dtype=np.int64)   # dataset.py:114
    src = keep[order]
    p = _Plan()
    p.keys = [keys_in[i] for i in src]
    p.labels = [labels_in[i] for i in src]
    p.src = src
    p.stage1 = stage1[src]
    p.stage2 = stage2[src]
    p.frames = frames[src].astype(np.int32)
    p.sample_rate = target_rate
    p.wav_dither = float(conf.get('wav_dither', 0.0) or 0.0)      # kaldi.fbank(dither=...), dataset.py:98
    # extension: 'sinc' (default) = torchaudio's width-6 hann sinc; 'kaiser' = the long sox-quality filter
    # (frontend.kaiser_sinc_kernel), for users who need the reference's libsox-grade speed perturb
    p.resampler = conf.get('resampler', 'sinc')
    return p


def _load_batch(batch):
    """Decodes every item; a failing item is reported like the reference does and marked not loaded."""
    waves, rates, loaded = [], [], []
    for x in batch:
        try:
            pcm, sr = _load_item(x)
            waves.append(pcm)
            rates.append(sr)
            loaded.append(True)
        except (Exception) as e:                                 # dataset.py:108-111
            print(e)
            logging.warning('read utterance {} error'.format(x[0]))
            waves.append(np.zeros(0, np.int16))
            rates.append(16000)
            loaded.append(False)
    return waves, rates, loaded


def _run_plan(plan, mel_bins, dev_wav, offs, lens, out_layout='padded', **fused):
    """The device half: [resample chain] -> fused fbank of every planned utterance into one tensor.
    ``dev_wav`` is the packed device buffer (int16 or fp32) the plan's ``src`` indices refer to."""
    fe = default_frontend(mel_bins, plan.sample_rate)
    B = len(plan.src)
    F = mel_bins
    if B == 0:
        return None, fe
    tmax = int(plan.frames.max())
    if out_layout == 'padded':
        out = torch.empty((B, tmax, F), dtype=torch.float32, device=fe.device)
        rows = np.arange(B, dtype=np.int64) * tmax
        nrows = np.full(B, tmax, dtype=np.int32)
    else:
        out = torch.empty((int(plan.frames.sum()), F), dtype=torch.float32, device=fe.device)
        rows = np.concatenate([[0], np.cumsum(plan.frames[:-1].astype(np.int64))]).astype(np.int64)
        nrows = plan.frames.copy()
    offs = np.asarray(offs, dtype=np.int64)
    lens = np.asarray(lens, dtype=np.int32)
    needs = (plan.stage1[:, 0] != 0) | (plan.stage2[:, 0] != 0)
    # speed 0.9 / 1.1 on int16 PCM: resampled inside the fbank kernel's staging -> the whole batch is ONE call
    if plan.wav_dither != 0.0:                               # Philox key per call; os.urandom: Python's `random` stream stays the reference's
        fused = dict(fused, wav_dither=plan.wav_dither,
                     dither_seed=fused.get('dither_seed') or int.from_bytes(os.urandom(8), 'little'))
    # (the library fuses only tables whose bits equal its baked torchaudio tables: ask it, fall back to oe_resample)
    ratio_keys = np.unique(plan.stage2[:, 0] * 65536 + plan.stage2[:, 1])
    fusable = (plan.wav_dither == 0.0 and dev_wav.dtype == torch.int16 and not (plan.stage1[:, 0] != 0).any() and
               plan.resampler == 'sinc' and fe.mel_baked and
               all(int(k) in (0, 9 * 65536 + 10, 11 * 65536 + 10) for k in ratio_keys) and
               all(fe.fusable(int(k) >> 16, int(k) & 65535) for k in ratio_keys if k))
    if fusable and needs.any():
        kw = dict(fused)
        if kw.get('frame_map') is not None:
            fm, starts = kw.pop('frame_map'), kw.pop('frame_map_starts')
            kw['frame_maps'] = [fm[starts[i]:starts[i] + plan.frames[i]] for i in range(B)]
        fe.fbank(dev_wav, offs[plan.src], lens[plan.src], layout='custom', out=out.view(-1, F), out_rows=rows,
                 out_nrows=nrows, speed_ratios=plan.stage2, **kw)
        return out, fe
    direct = np.nonzero(~needs)[0]
    resamp = np.nonzero(needs)[0]

    def call(wav, o, l, idx):
        kw = dict(fused)
        for k in ('tmask', 'fmask'):
            if kw.get(k) is not None:
                kw[k] = np.ascontiguousarray(kw[k][idx])
        if kw.get('frame_map') is not None:                      # concatenated map -> this subset's rows
            fm, starts = kw.pop('frame_map'), kw.pop('frame_map_starts')
            kw['frame_maps'] = [fm[starts[i]:starts[i] + plan.frames[i]] for i in idx]
        fe.fbank(wav, o, l, layout='custom', out=out.view(-1, F), out_rows=rows[idx], out_nrows=nrows[idx], **kw)

    if len(direct):
        call(dev_wav, offs[plan.src[direct]], lens[plan.src[direct]], direct)
    if len(resamp):
        cur, cur_offs, cur_lens = dev_wav, offs[plan.src[resamp]], lens[plan.src[resamp]]
        for stage, kind in ((plan.stage1[resamp], 'sinc'), (plan.stage2[resamp], plan.resampler)):   # resample_rate stage, then speed stage
            if (stage[:, 0] != 0).any():
                cur, cur_offs, cur_lens = fe.resample(cur, cur_offs, cur_lens, stage, kind=kind)
        call(cur, cur_offs, cur_lens, resamp)
    return out, fe


def _pack_loaded(waves):
    any_f32 = any(np.asarray(w).dtype.kind == 'f' for w in waves)
    return pack_waveforms(waves, dtype=np.float32 if any_f32 else np.int16)


def _extract_feature(batch, feature_extraction_conf):
    """openeat/dataset/dataset.py:39-118.  Returns (sorted_keys, sorted_feats, sorted_labels) with
    ``sorted_feats`` a list of (T_i, mel_bins) float32 numpy arrays, longest first."""
    waves, rates, loaded = _load_batch(batch)
    plan = _plan_batch([x[0] for x in batch], [np.array(x[2]) for x in batch], [len(w) for w in waves], rates,
                       [x[3] for x in batch], feature_extraction_conf, loaded=loaded)
    out = None
    if len(plan.src):
        buf, offs, lens = _pack_loaded(waves)
        fe = default_frontend(feature_extraction_conf['mel_bins'])
        out, _ = _run_plan(plan, feature_extraction_conf['mel_bins'], buf.to(fe.device, non_blocking=True), offs, lens,
                           out_layout='ragged')
    feats = []
    if out is not None:
        host = out.cpu().numpy()
        r = 0
        for m in plan.frames:
            feats.append(host[r:r + m])
            r += int(m)
    return plan.keys, feats, plan.labels


def _load_feature(batch):
    """Load acoustic features from Kaldi archives -- openeat/dataset/dataset.py:120-152.
    ``batch``: list of (key, 'file.ark:offset', tokenids, ...).  Returns (keys, feats, labels) sorted by
    length descending.  Two reference behaviours are kept on purpose: an utterance that fails to load is
    logged and dropped *after* its key was appended (dataset.py:136-146), and the label is appended twice
    per utterance (dataset.py:141,143), so ``sorted_labels[j] = labels[order[j]]`` indexes a doubled list."""
    keys, feats, lengths, labels = [], [], [], []
    for x in batch:
        try:
            keys.append(x[0])
            mat = kaldi_io.read_mat(x[1])
            feats.append(mat)
            lengths.append(mat.shape[0])
            labels.append(np.array(x[2]))
            labels.append(np.array(x[2]))
        except Exception:
            logging.warning('read utterance {} error'.format(x[0]))
    order = np.argsort(lengths)[::-1]
    sorted_keys = [keys[i] for i in order]
    sorted_feats = [feats[i] for i in order]
    sorted_labels = [labels[i] for i in order]
    return sorted_keys, sorted_feats, sorted_labels


class audio_collate_func(object):
    """Collate function for AudioDataset -- openeat/dataset/dataset.py:155-239, ``data_type='wav'``.

    Extra keyword-only options (not in the reference): ``output_device`` ('cuda' default, or 'cpu'),
    ``global_cmvn=(mean, istd)`` fp32 tensors to fuse GlobalCMVN (applied to padding too, exactly like
    the encoder does on the padded batch), ``cmvn_stats`` a float64 [2F+1] device tensor to accumulate
    raw-feature statistics into.
    """

    def __init__(self, feature_dither=0.0, spec_aug=False, spec_aug_conf=None, spec_sub=False,
                 spec_sub_conf=None, data_type="kaldi", feature_extraction_conf=None, normalization=True,
                 *, output_device='cuda', global_cmvn=None, cmvn_stats=None):
        self.feature_dither = feature_dither
        self.spec_sub = spec_sub
        self.spec_aug = spec_aug
        self.spec_sub_conf = spec_sub_conf
        self.spec_aug_conf = spec_aug_conf
        self.data_type = data_type
        self.feature_extraction_conf = feature_extraction_conf
        self.normalization = normalization
        self.output_device = output_device
        self.global_cmvn = global_cmvn
        self.cmvn_stats = cmvn_stats
        print('normalize feature', self.normalization)              # dataset.py:183
        # feature_dither (dataset.py:199-201): `a` is drawn with the reference's random.uniform call; the per-cell
        # uniforms come from Philox on the GPU (keyed by a per-object seed and a batch counter), not from numpy's
        # global generator, which is left untouched -- the option is stochastic, there is no value parity
        self._dither_seed = int.from_bytes(os.urandom(8), 'little')
        self._dither_batches = 0

    def __call__(self, batch):
        if len(batch) == 1:                                          # dataset.py:186-187
            batch = batch[0]
        if self.data_type != 'wav':                                  # dataset.py:190-191
            keys, xs, ys = _load_feature(batch)
            return self.collate_features(keys, xs, ys)
        if (len(batch) and torch.cuda.is_available() and os.environ.get('OE_FLAC_GPU', '1') != '0'
                and all(isinstance(x[1], str) and x[1].split(',')[0].strip().lower().endswith('.flac') for x in batch)):
            # FLAC lists (the LibriSpeech recipe): the compressed files cross PCIe, the GPU decodes them (oe_flac_gpu.cuh);
            # anything that decoder does not take (stereo, 24 bit, unreadable files) sends the batch to the host decoders below
            from .ingest import default_flac_ingest
            ing = default_flac_ingest()
            keys = [x[0] for x in batch]
            fb = ing.pack([x[1] for x in batch], keys, report=False)
            if fb.loaded.all():
                dev = fb.to_device(default_frontend(self.feature_extraction_conf['mel_bins']).device)
                ev = torch.cuda.Event()
                ev.record()
                ing.release_after(fb.slot, ev)
                lens = fb.drop_failed(fb.lens, fb.loaded, keys)
                return self.collate_packed(dev, fb.offsets, lens, keys, [x[2] for x in batch], [x[3] for x in batch],
                                           sample_rates=fb.rates, loaded=fb.loaded)
            ing.release_after(fb.slot, None)
        if len(batch) and all(isinstance(x[1], str) for x in batch):
            # wav files: native ingest (headers + multi-threaded pread straight into a pinned buffer, no per-utterance Python)
            from .ingest import default_ingest
            ing = default_ingest()
            keys = [x[0] for x in batch]
            buf, offs, lens, rates, loaded, slot = ing.load([x[1] for x in batch], keys, report=False)
            # the native ingest reads 16-bit PCM; a file it rejects but read_wav decodes (24-bit, float, ...: torchaudio.load
            # reads those too) sends the whole batch through the general Python decoder below
            if loaded.all() or not any(_decodable(batch[i][1]) for i in np.nonzero(~loaded)[0]):
                ing.report_errors(keys)
                out = self.collate_packed(buf, offs, lens, keys, [x[2] for x in batch], [x[3] for x in batch],
                                          sample_rates=rates, loaded=loaded)
                if torch.cuda.is_available():
                    ev = torch.cuda.Event()
                    ev.record()                  # behind the H2D copy of `buf`: the ring slot is reused after it
                    ing.release_after(slot, ev)
                return out
        waves, rates, loaded = _load_batch(batch)
        buf, offs, lens = _pack_loaded(waves)
        return self.collate_packed(buf, offs, lens, [x[0] for x in batch], [x[2] for x in batch],
                                   [x[3] for x in batch], sample_rates=rates, loaded=loaded)

    def collate_packed(self, wav, offsets, lens, keys, labels, speeds=None, sample_rates=None, loaded=None):
        """Same result as ``__call__`` for a batch that is already decoded and packed: ``wav`` is one int16
        (or fp32, int16 scale) tensor -- pinned host memory (copied asynchronously) or already on the
        device -- holding utterance i at ``offsets[i]`` (multiples of 8 samples) with ``lens[i]`` samples.
        This is the entry point a native data loader hands its PCM to."""
        conf = self.feature_extraction_conf
        B_in = len(keys)
        if B_in and not wav.is_cuda:             # start the H2D copy first: it overlaps the host-side planning
            wav = wav.to(default_frontend(conf['mel_bins']).device, non_blocking=True)
        speeds = [1.0] * B_in if speeds is None else speeds
        sample_rates = [16000] * B_in if sample_rates is None else sample_rates
        plan = _plan_batch(keys, labels, lens, sample_rates, speeds, conf, loaded=loaded)
        F = conf['mel_bins']
        frames = plan.frames
        fused = {'normalization': bool(self.normalization)}
        self._dither_a = random.uniform(0, self.feature_dither) if self.feature_dither != 0.0 else 0.0   # dataset.py:200
        # dataset.py:204-209: every spec_sub draw, then every spec_aug draw, in length-sorted order
        fmap, tmask, fmask = planner.plan_augment(frames, F, self.spec_sub_conf if self.spec_sub else None,
                                                  self.spec_aug_conf if self.spec_aug else None)
        if fmap is not None:
            fused['frame_map'] = fmap
            fused['frame_map_starts'] = np.concatenate([[0], np.cumsum(frames[:-1].astype(np.int64))]) if len(frames) else []
        if tmask is not None:
            fused['tmask'] = tmask
        if fmask is not None:
            fused['fmask'] = fmask
        if self.global_cmvn is not None:
            fused['cmvn'] = self.global_cmvn
            fused['cmvn_on_padding'] = True
        if self.cmvn_stats is not None:
            fused['stats'] = self.cmvn_stats
        fused.update(self._draw_dither())
        features = None
        if len(plan.src):
            features, _ = _run_plan(plan, F, wav, offsets, lens, out_layout=getattr(self, '_out_layout', 'padded'), **fused)
        return self._finish(plan.keys, features, frames, plan.labels)

    def _draw_dither(self):
        """dataset.py:199-201: one random.uniform per batch, placed between the speed draws (made while planning)
        and the spec_sub / spec_aug draws -- callers invoke this BEFORE planner.plan_augment."""
        if self.feature_dither == 0.0:
            return {}
        self._dither_batches += 1
        return {'feature_dither': self._dither_a, 'dither_seed': self._dither_seed + self._dither_batches}

    def collate_features(self, keys, xs, ys):
        """The post-fbank chain for features that already exist (``data_type != 'wav'``, dataset.py:190-238):
        ``xs`` is a list of (T_i, F) float arrays sorted by length as ``_load_feature`` returns them."""
        F = int(xs[0].shape[1]) if len(xs) else (self.feature_extraction_conf or {}).get('mel_bins', 80)
        frames = np.array([x.shape[0] for x in xs], dtype=np.int32)
        self._dither_a = random.uniform(0, self.feature_dither) if self.feature_dither != 0.0 else 0.0
        fmap, tmask, fmask = planner.plan_augment(frames, F, self.spec_sub_conf if self.spec_sub else None,
                                                  self.spec_aug_conf if self.spec_aug else None)
        features = None
        if len(xs):
            fe = default_frontend(F)
            offs = np.concatenate([[0], np.cumsum(frames[:-1].astype(np.int64))]).astype(np.int64)
            host = torch.from_numpy(np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float32) for x in xs])))
            kw = dict(normalization=bool(self.normalization), tmask=tmask, fmask=fmask, features_in=True)
            if fmap is not None:
                kw['frame_maps'] = [fmap[offs[i]:offs[i] + frames[i]] for i in range(len(xs))]
            if self.global_cmvn is not None:
                kw['cmvn'] = self.global_cmvn
                kw['cmvn_on_padding'] = True
            if self.cmvn_stats is not None:
                kw['stats'] = self.cmvn_stats
            kw.update(self._draw_dither())
            features, _ = fe.fbank(host.pin_memory().to(fe.device, non_blocking=True), offs, frames, layout='padded', **kw)
        return self._finish(keys, features, frames, ys)

    def _finish(self, keys, features, frames, ys):
        dev = torch.device(self.output_device)
        flen = np.array(frames, dtype=np.int32)
        tlen = np.fromiter(map(len, ys), dtype=np.int32, count=len(ys))
        if features is None:                                         # dataset.py:219-220
            features = torch.Tensor([])
            tpad = None
        else:                                                        # pad_sequence(..., True, IGNORE_ID), dataset.py:225-226
            tpad = np.full((len(ys), int(tlen.max()) if len(ys) else 0), IGNORE_ID, dtype=np.int32)
            if tpad.size:
                import itertools
                tpad[np.arange(tpad.shape[1])[None, :] < tlen[:, None]] = np.fromiter(
                    itertools.chain.from_iterable(ys), dtype=np.int32, count=int(tlen.sum()))
        if dev.type == 'cuda' and features.is_cuda:
            # the three int32 tensors travel as ONE block through the front-end's mapped pinned ring (oe_upload_small):
            # torch's .to() of a pageable tensor is a synchronous cudaMemcpy that queues behind the next batch's PCM
            fe = default_frontend((self.feature_extraction_conf or {}).get('mel_bins', features.shape[-1]))
            n, m = len(flen), (tpad.size if tpad is not None else 0)
            blob = fe.upload_small(np.concatenate([flen, tlen, tpad.reshape(-1)]) if m else np.concatenate([flen, tlen]))
            inputs = {'features': features, 'features_length': blob[:n], 'targets': blob[2 * n:].view(tpad.shape),
                      'targets_length': blob[n:2 * n]}
            return keys, inputs
        targets = torch.Tensor([]) if tpad is None else torch.from_numpy(tpad)
        inputs = {'features': features.to(dev), 'features_length': torch.from_numpy(flen).to(dev),
                  'targets': targets.to(dev), 'targets_length': torch.from_numpy(tlen).to(dev)}
        return keys, inputs


def _default_tokenizer(text):
    """Character / word tokens the way openeat/dataset/text_processor.py:2-22 splits WITHOUT a BPE model:
    every CJK character is a token, every other run of text one token.  Text normalisation and BPE are not
    part of the front-end path: pass the recipe's own callable as ``tokenizer`` for
    identical token ids."""
    import re
    pattern = re.compile(r'([\u4e00-\u9fff])')
    tokens = []
    for ch_or_w in [w for w in pattern.split(text.upper()) if len(w.strip()) > 0]:
        tokens.append(ch_or_w)
    return tokens


class AudioDataset(torch.utils.data.Dataset):
    """openeat/dataset/dataset.py:241-376: parses ``format.data`` (one utterance per line, tab-separated
    ``key:value`` fields ``utt feat feat_shape text`` [+ ``token tokenid token_shape``]; ``feat_shape`` is seconds for
    ``data_type='wav'`` and ``frames,dim`` for Kaldi-archive features, whose ``feat`` is ``file.ark:offset``),
    applies the length filters, the offline speed list, the optional sort and the static / dynamic / shuffle
    batching -- each item is a PRE-BUILT batch ``[(key, path, tokenid, speed), ...]`` for
    ``audio_collate_func`` -- including the reference's quirks (``num_frames *= speed``
    accumulates over the speed list and only affects sorting / batching; the 'dynamic' loop leaves an empty
    first batch when the first utterance alone exceeds ``max_frames_in_batch``).

    ``tokenizer``: callable text -> list of tokens applied to the ``text`` field of 4-field lines (default:
    the reference's CJK / non-CJK split without BPE or punctuation stripping)."""

    def __init__(self, data_file, char_dict, bpe_model=None, max_length=10240, min_length=0, token_max_length=200,
                 token_min_length=0, batch_type='static', batch_size=1, max_frames_in_batch=0, sort=False,
                 speed_perturb=False, speeds=[0.9, 1.1, 0.1], data_type="kaldi", tokenizer=None):
        import codecs
        assert batch_type in ['static', 'dynamic', 'shuffle']
        if bpe_model is not None and tokenizer is None:
            raise NotImplementedError('pass tokenizer= (e.g. a sentencepiece-backed callable); BPE is outside the front-end path')
        tokenizer = tokenizer or _default_tokenizer
        self.batch_size = 1 if batch_type in ['static', 'dynamic'] else batch_size       # dataset.py:295
        self.char_dict = char_dict
        self.vocab_size = len(char_dict)
        if speed_perturb:
            speed_list = [float(s) for s in np.arange(speeds[0], speeds[1], speeds[2])]   # dataset.py:298-301
        else:
            speed_list = [1.0]
        data = []
        with codecs.open(data_file, 'r', encoding='utf-8') as f:
            for line in f:
                arr = line.strip().split('\t')
                if len(arr) != 4 and len(arr) != 7:
                    continue
                key = arr[0].split(':')[1]
                if len(arr) == 4:
                    text = arr[3].split(':')[1]
                    tokens = tokenizer(text)
                    tokenid = [char_dict[w] if w in char_dict else char_dict['<unk>'] for w in tokens]
                else:
                    tokenid = arr[5].split(':')[1]                 # dataset.py:318-319 keeps the string
                path = ':'.join(arr[1].split(':')[1:])
                if data_type == 'wav':
                    num_frames = int(float(arr[2].split(':')[1]) * 1000 / 10)            # dataset.py:324
                else:                                              # dataset.py:325-331: feat_shape:<frames>,<dim>
                    feat_info = arr[2].split(':')[1].split(',')
                    feat_dim = int(feat_info[1].strip())
                    num_frames = int(feat_info[0].strip())
                    self.input_size = feat_dim
                length = num_frames
                token_length = len(tokenid)
                if min_length < length < max_length and token_min_length < token_length < token_max_length:
                    for speed in speed_list:
                        num_frames *= speed                        # dataset.py:334-336 (accumulates)
                        data.append((key, path, num_frames, tokenid, speed))
        if sort:
            data = sorted(data, key=lambda x: x[2])
        num_data = len(data)
        if batch_type == 'dynamic':                                # dataset.py:341-352
            assert (max_frames_in_batch > 0)
            self.data = [[]]
            num_frames_in_batch = 0
            for i in range(num_data):
                length = data[i][2]
                num_frames_in_batch += length
                if num_frames_in_batch > max_frames_in_batch:
                    self.data.append([])
                    num_frames_in_batch = length
                self.data[-1].append((data[i][0], data[i][1], data[i][3], data[i][4]))
        elif batch_type == 'static':                               # dataset.py:355-364
            self.data = []
            cur = 0
            while cur < num_data:
                end = min(cur + batch_size, num_data)
                self.data.append([(data[i][0], data[i][1], data[i][3], data[i][4]) for i in range(cur, end)])
                cur = end
        else:                                                      # dataset.py:365-368
            self.data = [[data[i][0], data[i][1], data[i][3], data[i][4]] for i in range(num_data)]
        print(len(self.data))                                      # dataset.py:370

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return self.data[idx]


class PrefetchingCollator(object):
    """Pipelines ``audio_collate_func.collate_packed`` over an iterator of packed host batches: the pinned
    PCM of batch i+1 crosses PCIe on a side stream while batch i runs its kernels, so the steady-state cost
    per batch is max(H2D copy, kernels, host planning) instead of their sum.  Yields exactly what
    ``collate_packed`` returns, in order (what torch's DataLoader prefetching does for the reference's
    CPU workers, done here for the one resource the GPU front-end is bound by: the H2D copy).

    ``batches`` yields tuples ``(pinned_wav, offsets, lens, keys, labels, speeds)``, optionally followed by
    ``sample_rates, loaded, release`` (``openeat_b200.ingest.ingest_batches``: batches read from wav files by the native
    ingest on a background thread; ``release(event)`` returns the pinned ring slot once the H2D copy has completed).

    ``to_host=True`` is the reference boundary proper (dataset.py:232-238: CPU tensors): the padded feature tensor
    of batch i goes back over PCIe on a third stream into a ring of pinned buffers while batch i+1 is being
    copied in and computed (PCIe is full duplex); a batch is handed out once its copy has landed, i.e. one
    batch later than it was launched.  The yielded ``features`` tensor is a view of a ring slot: it stays valid
    until ``ring`` more batches have been drawn.

    ``host_pad=True`` (with ``to_host``) brings back only the REAL rows (the kernels write the ragged layout: about half
    the bytes of the padded tensor for 2-10 s utterances) and lets a helper thread build the zero-padded
    ``(B, Tmax, F)`` tensor on the host (``oe_host_pad_rows``: reader-pool threads, non-temporal stores; padding rows are
    0, or ``(0 - mean) * istd`` with a fused GlobalCMVN); one more batch of latency.  It trades PCIe bytes for host
    memory traffic: on the measured 16-core box, whose memory the two copy engines already keep busy, it is SLOWER
    (0.69 M against 0.85 M audio-s/s) -- an option for hosts with a narrow link, not the default.
    """

    def __init__(self, collate, batches, to_host=False, ring=3, host_pad=False):
        self.collate = collate
        self.batches = iter(batches)
        fe = default_frontend(collate.feature_extraction_conf['mel_bins'])
        self.device = fe.device
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.to_host = bool(to_host)
        self.host_pad = bool(to_host and host_pad)
        self.out_stream = torch.cuda.Stream(device=self.device) if to_host else None
        self._depth = 2 if self.host_pad else 1  # batches in flight behind the one being launched
        self._ring = [None] * (max(2, int(ring)) + self._depth)
        self._slot = 0
        self._stage_ring = [None] * (self._depth + 2)   # pinned landing buffers of the ragged rows
        self._stage_slot = 0
        self._pending = []                       # launched batches, oldest first: (keys, host dict, event, thread or None)
        self._next = None
        self._staged = collections.deque()
        self._stage_depth = 1
        self._exhausted = False
        self._pad_row = None
        if self.host_pad:
            import threading
            from .ingest import NativeIngest
            self._lock = threading.Lock()        # the helper threads of consecutive batches take ring slots
            self._padder = NativeIngest(threads=int(os.environ.get('OE_PAD_THREADS', '0')), ring=2)
            if getattr(collate, 'global_cmvn', None) is not None:
                mean, istd = collate.global_cmvn
                m = mean.detach().float().cpu().numpy()
                self._pad_row = (np.float32(0.0) - m) if istd is None else (np.float32(0.0) - m) * istd.detach().float().cpu().numpy()
        self._stage()

    def _stage(self):
        """Keeps ``_stage_depth`` batches staged (copies enqueued) ahead of the one being launched: 1 for PCM; 2 for FLAC
        batches, whose H2D copy (copy stream) and decode kernel (decode stream) then overlap the previous batch's."""
        while not self._exhausted and len(self._staged) < self._stage_depth:
            try:
                item = next(self.batches)
            except StopIteration:
                self._exhausted = True
                break
            wav = item[0]
            flac = wav if hasattr(wav, 'to_device') else None
            with torch.cuda.stream(self.copy_stream):
                if flac is not None:             # ingest.FlacBatch: compressed bytes cross PCIe, the decode kernel follows on its own stream
                    self._stage_depth = 2
                    dev = flac.to_device(self.device, wait=False)
                    ev = flac.event
                else:
                    dev = wav if wav.is_cuda else wav.to(self.device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.copy_stream)
            if len(item) > 8 and item[8] is not None:
                item[8](ev)                      # the ingest ring slot is free once this copy has completed
            self._staged.append(((dev, ev) + tuple(item[1:8]), flac))
        self._next = self._staged[0][0] if self._staged else None

    def __iter__(self):
        return self

    def _launch(self):
        (dev, ev, offs, lens, keys, labels, speeds), flac = self._staged[0][0][:7], self._staged[0][1]
        extra = self._staged[0][0][7:]
        self._staged.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)                       # kernels of this batch start when its PCM has landed
        dev.record_stream(cur)
        if flac is not None and len(extra) >= 2:
            # streams that failed the GPU decoder's checks are dropped like any unreadable file (dataset.py:108-111); the
            # batch was staged two steps ahead, so its decode has normally finished by now
            lens = flac.drop_failed(lens, extra[1], keys)
        self._stage()                            # next batch's copy is in flight before this batch's kernels are enqueued
        self.collate._out_layout = 'ragged' if self.host_pad else 'padded'
        try:
            if len(extra) >= 2:
                return self.collate.collate_packed(dev, offs, lens, keys, labels, speeds, sample_rates=extra[0], loaded=extra[1])
            return self.collate.collate_packed(dev, offs, lens, keys, labels, speeds)
        finally:
            self.collate._out_layout = 'padded'

    def _to_host(self, keys, inputs):
        """Enqueues the D2H copies of one finished batch on the output stream (and, with host padding, starts the thread
        that completes the padded tensor); returns (keys, host dict, event, thread)."""
        cur = torch.cuda.current_stream(self.device)
        done = torch.cuda.Event()
        done.record(cur)
        host = {}
        ragged = None
        with torch.cuda.stream(self.out_stream):
            self.out_stream.wait_event(done)
            for k, v in inputs.items():
                if not v.is_cuda:
                    host[k] = v
                    continue
                if k == 'features' and v.numel():
                    if self.host_pad:
                        slot = self._stage_ring[self._stage_slot]
                        if slot is None or slot.numel() < v.numel():
                            slot = torch.empty(int(v.numel() * 1.2) + 1024, dtype=v.dtype).pin_memory()
                            self._stage_ring[self._stage_slot] = slot
                        self._stage_slot = (self._stage_slot + 1) % len(self._stage_ring)
                        dst = ragged = slot[:v.numel()].view(v.shape)
                    else:
                        dst = self._out_slot(v.numel(), v.dtype).view(v.shape)
                else:
                    dst = torch.empty(v.shape, dtype=v.dtype).pin_memory() if v.numel() else torch.empty(v.shape, dtype=v.dtype)
                dst.copy_(v, non_blocking=True)
                v.record_stream(self.out_stream)
                host[k] = dst
            ev = torch.cuda.Event()
            ev.record(self.out_stream)
        thread = None
        if ragged is not None:
            import threading

            def finish():
                ev.synchronize()                     # the rows and the frame counts have landed (GIL released while waiting)
                frames = host['features_length'].numpy()
                tmax, F = int(frames.max()), int(ragged.shape[-1])
                with self._lock:
                    out = self._out_slot(len(frames) * tmax * F, ragged.dtype).view(len(frames), tmax, F)
                    self._padder.pad_rows(ragged, frames, tmax, out, self._pad_row)
                host['features'] = out
            thread = threading.Thread(target=finish, daemon=True)
            thread.start()
        return keys, host, ev, thread

    def _out_slot(self, numel, dtype):
        slot = self._ring[self._slot]
        if slot is None or slot.numel() < numel or slot.dtype != dtype:
            slot = torch.empty(int(numel * 1.2) + 1024, dtype=dtype).pin_memory()
            self._ring[self._slot] = slot
        self._slot = (self._slot + 1) % len(self._ring)
        return slot[:numel]

    def __next__(self):
        if not self.to_host:
            if self._next is None:
                raise StopIteration
            return self._launch()
        if self._next is None and not self._pending:
            raise StopIteration
        # keep `_depth` batches in flight behind the one handed out
        while self._next is not None and len(self._pending) <= self._depth:
            self._pending.append(self._to_host(*self._launch()))
        keys, host, ev, thread = self._pending.pop(0)
        if thread is not None:
            thread.join()
        else:
            ev.synchronize()
        return keys, host
