"""Host-side driver of the CUDA front-end: owns the handle, packs ragged batches, launches.

PyTorch is used here only for device memory, streams and pinned staging; all numeric work
happens in ``csrc/oe_frontend.cu`` through the C ABI (``include/openeat_frontend.h``).
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import (OE_FEATS_F32, OE_NORM_NONE, OE_NORM_PER_UTT, OE_RS_DIRECT, OE_WAV_F32, OE_WAV_I16, FrontendError,
                   OeBatch, OeConfig, OeResampleBatch, c_f32p, c_i32p, c_i64p, check)

ALIGN = 8  # samples; the kernels read 16-byte vectors (include/openeat_frontend.h: wav_offsets)


def torch_povey_window(n=400):
    """kaldi.py:98-100 with torch's own fp32 ops -> bit-identical to torchaudio's table."""
    return torch.hann_window(n, periodic=False, dtype=torch.float32).pow(0.85)


def torch_mel_banks(num_bins=80, padded=512, sample_freq=16000.0, low_freq=20.0, high_freq=0.0):
    """kaldi.py:436-511 (vtln_warp_factor == 1.0) with torch's own fp32 ops."""
    num_fft_bins = padded / 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return torch.max(torch.zeros(1), torch.min(up, down))


def torch_sinc_kernel(orig, new, lowpass_filter_width=6, rolloff=0.99):
    """torchaudio functional.py:1343-1398 (sinc_interp_hann) on an fp32 grid, as the reference's
    substitute speed oracle evaluates it for fp32 waveforms.  Returns (kernel[new, taps], width)."""
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float32)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float32)[:, None, None] / new + idx
    t *= base
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels[:, 0, :].contiguous(), width


def kaiser_sinc_kernel(orig, new, passband=0.95, rejection_db=125.0):
    """Polyphase table of a long linear-phase low-pass with the specification of ``sox rate`` at its default quality
    (-h: 95 % of the band kept, 125 dB rejection, stop band from the lower Nyquist frequency): Kaiser-windowed sinc,
    designed in float64, in the layout ``oe_add_resampler`` takes (kernel[new, 2*width+orig]).  libsox itself (the
    reference's ``_speed_perturb``, audio_processor.py:31-34) is not reproducible here; this is the closest resampler the
    front-end offers to it (``resampler='kaiser'``), at ~25x the arithmetic of the default width-6 hann sinc."""
    fn = 0.5 * min(1.0, new / orig)                 # cycles per input sample
    f_pass, f_stop = passband * fn, fn
    beta = 0.1102 * (rejection_db - 8.7)
    half = int(math.ceil((rejection_db - 7.95) / (14.36 * (f_stop - f_pass)) / 2.0))
    fc = 0.5 * (f_pass + f_stop)
    q = np.arange(-half, half + orig, dtype=np.float64)[None, :]
    p = np.arange(new, dtype=np.float64)[:, None]
    t = q - p * orig / new
    with np.errstate(invalid='ignore', divide='ignore'):
        h = np.where(t == 0, 2.0 * fc, np.sin(2.0 * math.pi * fc * t) / (math.pi * t))
    r = np.clip(1.0 - (t / half) ** 2, 0.0, None)
    w = np.where(np.abs(t) <= half, np.i0(beta * np.sqrt(r)) / np.i0(beta), 0.0)
    return np.ascontiguousarray((h * w).astype(np.float32)), half


def speed_ratio(speed, sample_rate=16000):
    """torchaudio functional.py:2408-2413: 0.9 -> (9, 10), 1.1 -> (11, 10)."""
    src, dst = int(speed * sample_rate), int(sample_rate)
    g = math.gcd(src, dst)
    return src // g, dst // g


def aligned_offsets(lens, align=ALIGN):
    """Packed sample offsets with every utterance starting on an `align`-sample boundary."""
    lens = np.asarray(lens, dtype=np.int64)
    padded = (lens + align - 1) // align * align
    offs = np.zeros(len(lens), dtype=np.int64)
    if len(lens) > 1:
        offs[1:] = np.cumsum(padded[:-1])
    total = int(padded.sum())
    return offs, total


def pack_waveforms(waves, dtype=np.int16, pinned=True):
    """list of 1-D arrays -> (pinned packed host tensor, offsets, lens)."""
    lens = np.array([len(w) for w in waves], dtype=np.int32)
    offs, total = aligned_offsets(lens)
    tdt = torch.int16 if dtype == np.int16 else torch.float32
    buf = torch.zeros(max(total, ALIGN), dtype=tdt)
    if pinned and torch.cuda.is_available():
        buf = buf.pin_memory()
    view = buf.numpy()
    for w, o in zip(waves, offs):
        view[o:o + len(w)] = w
    return buf, offs, lens


def _ptr(arr, ptype):
    return arr.ctypes.data_as(ptype) if arr is not None else None


class PreparedBatch(object):
    """A batch description validated and packed by the library (``oe_batch_prepare``); see ``Frontend.prepare``."""

    def __init__(self, fe, handle, batch, out_shape, needs_out, keep):
        self.fe, self.handle, self.batch = fe, handle, batch
        self.out_shape, self.needs_out, self._keep = out_shape, needs_out, keep
        self.ws_bytes = int(fe.lib.oe_prepared_workspace_bytes(handle))
        fr = fe.lib.oe_prepared_frames(handle)
        self.frames = np.ctypeslib.as_array(fr, shape=(batch,)).copy() if batch else np.zeros(0, np.int32)

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            self.fe.lib.oe_prepared_destroy(h)
            self.handle = None


class Frontend(object):
    """One CUDA front-end handle (tables + resampler taps) on one device."""

    def __init__(self, mel_bins=80, sample_rate=16000, device=None, torch_tables=True, mel=None):
        """``torch_tables``: build the Povey window and the mel matrix with torch's own fp32 expressions (bit-identical to
        torchaudio's); False leaves both to the C library (its standard 80-bin table is the same matrix).  ``mel``: a
        custom (mel_bins, fft_size / 2) fp32 matrix instead."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise FrontendError('openeat_b200 needs a CUDA device: the front-end has no CPU path')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        cfg = OeConfig()
        check(self.lib.oe_config_default(ctypes.byref(cfg)))
        cfg.sample_rate = int(sample_rate)
        cfg.frame_length = int(sample_rate * 25.0 * 0.001)     # kaldi.py:138-140
        cfg.frame_shift = int(sample_rate * 10.0 * 0.001)
        cfg.fft_size = 1 if cfg.frame_length == 0 else 2 ** (cfg.frame_length - 1).bit_length()
        cfg.num_mel_bins = int(mel_bins)
        self.cfg = cfg
        self.mel_bins = int(mel_bins)
        self.sample_rate = int(sample_rate)
        custom_mel, win, mel = mel, None, None
        if torch_tables:
            win = np.ascontiguousarray(torch_povey_window(cfg.frame_length).numpy())
            mel = np.ascontiguousarray(torch_mel_banks(mel_bins, cfg.fft_size, float(sample_rate)).numpy())
        if custom_mel is not None:
            mel = np.ascontiguousarray(custom_mel, dtype=np.float32)
            if mel.shape != (int(mel_bins), cfg.fft_size // 2):
                raise FrontendError('mel must be (%d, %d)' % (mel_bins, cfg.fft_size // 2))
        handle = ctypes.c_void_p()
        check(self.lib.oe_frontend_create(ctypes.byref(cfg), _ptr(win, c_f32p), _ptr(mel, c_f32p),
                                          self.device.index, ctypes.byref(handle)))
        self.handle = handle
        self.torch_tables = torch_tables
        self._resamplers = {}
        self._fusable = {}
        self._mel_baked = None
        self._ws = {}
        self._extra_launches = 0     # oe_cmvn_apply takes no handle: counted here

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            self.lib.oe_frontend_destroy(h)
            self.handle = None

    # ------------------------------------------------------------------ helpers
    def tables(self):
        win = np.zeros(self.cfg.frame_length, np.float32)
        mel = np.zeros((self.mel_bins, self.cfg.fft_size // 2), np.float32)
        check(self.lib.oe_frontend_get_tables(self.handle, _ptr(win, c_f32p), _ptr(mel, c_f32p)))
        return win, mel

    def num_frames(self, n):
        return int(self.lib.oe_num_frames(self.handle, int(n)))

    def num_frames_array(self, lens):
        """kaldi.py:63-67 vectorised: 1 + (n - 400) // 160, 0 below one window."""
        lens = np.asarray(lens, dtype=np.int64)
        win, shift = self.cfg.frame_length, self.cfg.frame_shift
        return np.where(lens < win, 0, 1 + (lens - win) // shift).astype(np.int32)

    def _stream(self, stream):
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        return s, ctypes.c_void_p(s.cuda_stream)

    def _workspace(self, key, nbytes):
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ speed perturb
    def resampler_id(self, orig, new, kind='sinc'):
        """Table id of the (orig, new) polyphase resampler.  kind 'sinc': torchaudio's width-6 hann sinc (the substitute
        the path is pinned to); 'kaiser': the sox-quality design of ``kaiser_sinc_kernel``.  Returns OE_RS_DIRECT for a
        'sinc' ratio too long to tabulate (evaluated on the fly by ``resample``)."""
        key = (int(orig), int(new), kind)
        if key not in self._resamplers:
            tid = ctypes.c_int32(-1)
            if kind == 'kaiser':
                k, _ = kaiser_sinc_kernel(key[0], key[1])
                check(self.lib.oe_add_resampler(self.handle, key[0], key[1], _ptr(k, c_f32p), k.shape[1], ctypes.byref(tid)))
            elif kind != 'sinc':
                raise ValueError(kind)
            elif key[1] * (2 * math.ceil(6 * key[0] / (min(key[0], key[1]) * 0.99)) + key[0]) > 65536:
                tid = ctypes.c_int32(OE_RS_DIRECT)
            elif self.torch_tables:
                k, _ = torch_sinc_kernel(key[0], key[1])
                k = np.ascontiguousarray(k.numpy())
                check(self.lib.oe_add_resampler(self.handle, key[0], key[1], _ptr(k, c_f32p), k.shape[1],
                                                ctypes.byref(tid)))
            else:
                check(self.lib.oe_add_resampler(self.handle, key[0], key[1], None, 0, ctypes.byref(tid)))
            self._resamplers[key] = tid.value
        return self._resamplers[key]

    def fusable(self, orig, new):
        """True when the (orig, new) 'sinc' resampler can run inside the fbank kernel's staging (speed_ratios)."""
        key = (int(orig), int(new))
        if key not in self._fusable:
            tid = self.resampler_id(orig, new)
            self._fusable[key] = tid >= 0 and bool(self.lib.oe_resampler_fusable(self.handle, tid))
        return self._fusable[key]

    @property
    def mel_baked(self):
        if self._mel_baked is None:
            self._mel_baked = bool(self.lib.oe_mel_is_baked(self.handle))
        return self._mel_baked

    def resample_out_len(self, n, orig, new):
        return int(self.lib.oe_resample_out_len(int(n), int(orig), int(new)))

    def resample(self, wav, offsets, lens, ratios, out=None, out_offsets=None, stream=None, kind='sinc'):
        """Ragged polyphase resampling.  ``ratios[b]`` is (orig, new) or None (plain copy to fp32).
        Returns (fp32 device tensor, out_offsets, out_lens)."""
        B = len(lens)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        if not isinstance(ratios, np.ndarray):                   # list of (orig, new) / None -> [B, 2], (0, 0) = plain copy
            ratios = np.array([(0, 0) if r is None else tuple(r) for r in ratios], dtype=np.int64).reshape(B, 2)
        ratios = np.asarray(ratios, dtype=np.int64)
        ids = np.full(B, -1, dtype=np.int32)
        olens = lens.astype(np.int64)
        key = ratios[:, 0] * (1 << 32) + ratios[:, 1]
        for k in np.unique(key[key != 0]):
            o, n = int(k) >> 32, int(k) & 0xFFFFFFFF
            sel = key == k
            ids[sel] = self.resampler_id(o, n, kind)
            olens[sel] = (n * olens[sel] + o - 1) // o
        olens = olens.astype(np.int32)
        direct = ids == OE_RS_DIRECT
        r_orig = np.ascontiguousarray(ratios[:, 0], dtype=np.int32) if direct.any() else None
        r_new = np.ascontiguousarray(ratios[:, 1], dtype=np.int32) if direct.any() else None
        if out_offsets is None:
            out_offsets, total = aligned_offsets(olens)
        else:
            out_offsets = np.ascontiguousarray(out_offsets, dtype=np.int64)
            total = int((out_offsets + olens).max()) if B else 0
        if out is None:
            out = torch.empty(max(total, ALIGN), dtype=torch.float32, device=self.device)   # gaps are never read
        rb = OeResampleBatch(B, OE_WAV_F32 if wav.dtype == torch.float32 else OE_WAV_I16, _ptr(offsets, c_i64p),
                             _ptr(lens, c_i32p), _ptr(ids, c_i32p), _ptr(out_offsets, c_i64p), None,
                             _ptr(r_orig, c_i32p), _ptr(r_new, c_i32p))
        need = ctypes.c_size_t()
        check(self.lib.oe_resample_workspace_bytes(self.handle, ctypes.byref(rb), ctypes.byref(need)))
        s, sp = self._stream(stream)
        ws = self._workspace(('rs', s.cuda_stream), need.value)
        check(self.lib.oe_resample(self.handle, ctypes.byref(rb), ctypes.c_void_p(wav.data_ptr()),
                                   ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ws.data_ptr()),
                                   ws.numel(), sp))
        return out, out_offsets, olens

    # ------------------------------------------------------------------ fbank (+ fused chain)
    def fbank(self, wav, offsets, lens, *, layout='padded', out=None, max_rows=None, out_rows=None,
              out_nrows=None, normalization=False, tmask=None, fmask=None, frame_maps=None, cmvn=None,
              cmvn_on_padding=False, stats=None, stream=None, features_in=False, want_out=True,
              speed_ratios=None, feature_dither=0.0, dither_seed=0, wav_dither=0.0):
        """Runs one ragged batch.

        wav       packed device tensor: int16 PCM, fp32 on the int16 scale, or (features_in) a
                  (rows, mel_bins) fp32 feature matrix.
        offsets   [B] sample (row) offsets, lens [B] samples (rows).
        layout    'padded' -> (B, Tmax, F) zero padded like pad_sequence (dataset.py:217-218);
                  'ragged' -> (sum T_b, F);
                  'custom' -> caller gives out (rows, F), out_rows [B] and optionally out_nrows [B]
                  (used to let several calls fill one padded tensor).
        tmask / fmask   int32 [B, n, 2] half-open ranges (spec_aug plan) or None.
        frame_maps      list of int32 index maps (spec_sub plan) or None.
        cmvn      (mean, istd) fp32 device tensors, istd may be None (norm_var=False).
        stats     float64 device tensor [2F+1] accumulating sum / sumsq / count of raw frames.
        speed_ratios  int [B, 2] (orig, new) per utterance, (0, 0) = none: speed perturb FUSED into the
                  fbank kernel's staging (int16 input; 9:10 and 11:10 only, i.e. speeds 0.9 / 1.1).
        feature_dither  amplitude a of dataset.py:199-201 (x + (U[0,1) - 0.5) * a after the normalisation, before
                  spec_sub / spec_aug), 0 = off; uniforms from Philox keyed by dither_seed (no value parity).
        wav_dither      kaldi.fbank's dither (kaldi.py:179-181): + wav_dither * N(0,1) on every frame element, Philox +
                  Box-Muller keyed by dither_seed (no value parity); not with speed_ratios (resample first).
        Returns (out tensor or None, frames int32 ndarray).
        """
        if features_in:
            dtype = OE_FEATS_F32
        else:
            if wav.dtype not in (torch.float32, torch.int16):
                raise FrontendError('waveform must be int16 or float32, got %s' % wav.dtype)
            dtype = OE_WAV_F32 if wav.dtype == torch.float32 else OE_WAV_I16
        prep = self.prepare(dtype, offsets, lens, layout=layout, out_shape=None if out is None else tuple(out.shape),
                            max_rows=max_rows, out_rows=out_rows, out_nrows=out_nrows, normalization=normalization,
                            tmask=tmask, fmask=fmask, frame_maps=frame_maps, cmvn=cmvn,
                            cmvn_on_padding=cmvn_on_padding, stats=stats, want_out=want_out,
                            speed_ratios=speed_ratios, feature_dither=feature_dither, dither_seed=dither_seed,
                            wav_dither=wav_dither)
        out = self.run(prep, wav, out=out, stream=stream)
        return out, prep.frames

    def prepare(self, wav_dtype, offsets, lens, *, layout='padded', out_shape=None, max_rows=None, out_rows=None,
                out_nrows=None, normalization=False, tmask=None, fmask=None, frame_maps=None, cmvn=None,
                cmvn_on_padding=False, stats=None, want_out=True, speed_ratios=None, feature_dither=0.0,
                dither_seed=0, wav_dither=0.0):
        """Validates and packs one batch description natively (``oe_batch_prepare``) and returns a ``PreparedBatch``
        that ``run`` launches with no further host work -- same keywords as ``fbank`` (``wav_dtype``: OE_WAV_I16 /
        OE_WAV_F32 / OE_FEATS_F32 or a torch dtype; ``out_shape``: shape of a caller-provided output tensor).  A
        prepared batch can be run many times and on any stream; cmvn / stats tensors must outlive it."""
        if isinstance(wav_dtype, torch.dtype):
            wav_dtype = OE_WAV_F32 if wav_dtype == torch.float32 else OE_WAV_I16
        features_in = wav_dtype == OE_FEATS_F32
        B = len(lens)
        F = self.mel_bins
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        rs_ids = None
        if features_in:
            frames = lens.copy()
        else:
            eff = lens.astype(np.int64)
            if speed_ratios is not None:
                speed_ratios = np.asarray(speed_ratios, dtype=np.int64).reshape(B, 2)
                rs_ids = np.full(B, -1, dtype=np.int32)
                key = speed_ratios[:, 0] * 65536 + speed_ratios[:, 1]
                for k in np.unique(key[key != 0]):
                    o, n = int(k) >> 16, int(k) & 65535
                    sel = key == k
                    rs_ids[sel] = self.resampler_id(o, n)
                    eff[sel] = (n * eff[sel] + o - 1) // o
            frames = self.num_frames_array(eff)
        nrows = None
        shape = None
        if not want_out:
            out_rows = np.zeros(B, dtype=np.int64)
        elif layout == 'padded':
            tmax = int(frames.max()) if B else 0
            if max_rows is not None:
                tmax = max(tmax, int(max_rows))
            if out_shape is not None:
                tmax = out_shape[1]
            shape = (B, tmax, F)
            out_rows = np.arange(B, dtype=np.int64) * tmax
            nrows = np.full(B, tmax, dtype=np.int32)
        elif layout == 'ragged':
            if out_rows is None:
                out_rows = np.zeros(B, dtype=np.int64)
                if B > 1:
                    out_rows[1:] = np.cumsum(frames[:-1].astype(np.int64))
            else:
                out_rows = np.ascontiguousarray(out_rows, dtype=np.int64)
            shape = (int(frames.sum()), F)
        elif layout == 'custom':
            if out_shape is None or out_rows is None:
                raise FrontendError("layout='custom' needs out and out_rows")
            out_rows = np.ascontiguousarray(out_rows, dtype=np.int64)
            if out_nrows is not None:
                nrows = np.ascontiguousarray(out_nrows, dtype=np.int32)
        else:
            raise ValueError(layout)
        tm = fm = None
        n_t = n_f = 0
        if tmask is not None and np.size(tmask):
            tm = np.ascontiguousarray(tmask, dtype=np.int32).reshape(B, -1, 2)
            n_t = tm.shape[1]
        if fmask is not None and np.size(fmask):
            fm = np.ascontiguousarray(fmask, dtype=np.int32).reshape(B, -1, 2)
            n_f = fm.shape[1]
        fmap = fmap_off = None
        if frame_maps is not None:
            fmap_off = np.zeros(B, dtype=np.int64)
            if B > 1:
                fmap_off[1:] = np.cumsum(frames[:-1].astype(np.int64))
            if isinstance(frame_maps, np.ndarray) and frame_maps.ndim == 1:      # already concatenated in batch order
                fmap = np.ascontiguousarray(frame_maps, dtype=np.int32)
            else:
                fmap = np.concatenate([np.asarray(m, dtype=np.int32) for m in frame_maps]) if B else np.zeros(0, np.int32)
                fmap = np.ascontiguousarray(fmap, dtype=np.int32)
            if fmap.shape[0] != int(frames.sum()):
                raise FrontendError('frame_maps must have one entry per frame')
        mean_p = istd_p = None
        if cmvn is not None:
            mean, istd = cmvn
            mean_p = ctypes.c_void_p(mean.data_ptr())
            istd_p = ctypes.c_void_p(istd.data_ptr()) if istd is not None else None
        bt = OeBatch(B, wav_dtype, _ptr(offsets, c_i64p), _ptr(lens, c_i32p), _ptr(out_rows, c_i64p),
                     _ptr(nrows, c_i32p), 0, OE_NORM_PER_UTT if normalization else OE_NORM_NONE, n_t, n_f,
                     _ptr(tm, c_i32p), _ptr(fm, c_i32p), _ptr(fmap, c_i32p), _ptr(fmap_off, c_i64p),
                     mean_p, istd_p, 1 if cmvn_on_padding else 0,
                     ctypes.c_void_p(stats.data_ptr()) if stats is not None else None,
                     None, _ptr(rs_ids, c_i32p), float(feature_dither), int(dither_seed) & 0xFFFFFFFFFFFFFFFF,
                     float(wav_dither))
        handle = ctypes.c_void_p()
        check(self.lib.oe_batch_prepare(self.handle, ctypes.byref(bt), ctypes.byref(handle)))
        return PreparedBatch(self, handle, B, shape if want_out else None, layout == 'custom' and want_out,
                             (cmvn, stats))

    def run(self, prep, wav, out=None, stream=None):
        """Launches a prepared batch on ``wav`` (packed device tensor).  Returns the output tensor (None when the batch
        was prepared with want_out=False)."""
        if out is None and prep.out_shape is not None:
            out = torch.empty(prep.out_shape, dtype=torch.float32, device=self.device)
        if prep.needs_out and out is None:
            raise FrontendError("layout='custom' needs out")
        s, sp = self._stream(stream)
        ws = self._workspace(('fb', s.cuda_stream), prep.ws_bytes)
        check(self.lib.oe_fbank_run(self.handle, prep.handle, ctypes.c_void_p(wav.data_ptr()),
                                    ctypes.c_void_p(out.data_ptr()) if out is not None else None,
                                    ctypes.c_void_p(ws.data_ptr()), ws.numel(), sp))
        return out

    def upload_small(self, arr, stream=None):
        """Small numpy array -> new device tensor without the copy engine (``oe_upload_small``): the transfer does not
        queue behind a bulk H2D copy and the host never blocks."""
        arr = np.ascontiguousarray(arr)
        tdt = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32}[arr.dtype]
        n = arr.size
        pad = (-n * arr.itemsize) % 16 // arr.itemsize
        out = torch.empty(n + pad, dtype=tdt, device=self.device)
        if n:
            _, sp = self._stream(stream)
            check(self.lib.oe_upload_small(self.handle, ctypes.c_void_p(arr.ctypes.data), arr.nbytes,
                                           ctypes.c_void_p(out.data_ptr()), sp))
        return out[:n].view(arr.shape)

    def set_kernel_timing(self, on=True):
        """Measurement hook: bracket the fbank kernel of every following ``fbank`` call with CUDA events."""
        check(self.lib.oe_frontend_set_kernel_timing(self.handle, 1 if on else 0))

    def fbank_kernel_ms(self):
        """Duration of the fbank kernel of the most recent ``fbank`` call (waits for it); needs set_kernel_timing."""
        ms = ctypes.c_float()
        check(self.lib.oe_frontend_fbank_kernel_ms(self.handle, ctypes.byref(ms)))
        return float(ms.value)

    def step_ms(self):
        """Duration of the whole launch sequence of the most recent call (waits for it); needs set_kernel_timing."""
        ms = ctypes.c_float()
        check(self.lib.oe_frontend_step_ms(self.handle, ctypes.byref(ms)))
        return float(ms.value)

    @property
    def launches(self):
        """Kernels launched through this handle, counted by the library at every launch site
        (bench.py reports the difference over its timed region as gpu_launches)."""
        return int(self.lib.oe_frontend_launch_count(self.handle)) + self._extra_launches

    def cmvn_apply(self, x, mean, istd=None, out=None, stream=None):
        """GlobalCMVN.forward on a contiguous fp32 device tensor (..., F)."""
        x = x.contiguous()
        if out is None:
            out = torch.empty_like(x)
        _, sp = self._stream(stream)
        self._extra_launches += 1 if x.numel() else 0
        check(self.lib.oe_cmvn_apply(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                     x.numel() // x.shape[-1], x.shape[-1], ctypes.c_void_p(mean.data_ptr()),
                                     ctypes.c_void_p(istd.data_ptr()) if istd is not None else None, sp))
        return out


_default = {}


def default_frontend(mel_bins=80, sample_rate=16000, device=None):
    """Process-wide cache of handles keyed by (mel_bins, sample_rate, device)."""
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    key = (int(mel_bins), int(sample_rate), dev)
    if key not in _default:
        _default[key] = Frontend(mel_bins, sample_rate, torch.device('cuda', dev))
    return _default[key]
