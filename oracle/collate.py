"""CPU port of the reference's collate path (oracle; tests and the CPU baseline only).

Restates ``_extract_feature`` (``openeat/dataset/dataset.py:39-118``) and
``audio_collate_func`` (``dataset.py:155-239``) for ``data_type='wav'`` with the
same control flow, the same Python-``random`` call order and
the same drop-on-error convention.  Differences, all forced by this image:

* wav decoding uses the stdlib ``wave`` module (16-bit PCM) instead of libsox via
  ``torchaudio.load`` (no backend available).  ``torchaudio.load`` yields
  ``int16/32768`` as fp32 and the reference multiplies by ``1<<15``
  (dataset.py:75), i.e. the exact int16 values as fp32 -- which is what is
  produced here directly.  Items may also carry an in-memory int16 / float32
  array in place of the path (used by the benchmarks).
* speed perturb uses the substitute oracle (``oracle.speed``) instead of libsox.
* ``fbank_fn`` defaults to ``torchaudio.compliance.kaldi.fbank`` -- the very
  function the reference calls -- when torchaudio is importable, else to the
  numpy restatement ``oracle.fbank.fbank``.
"""
import logging
import random
import wave

import numpy as np

from . import augment, fbank as ofbank, speed as ospeed

IGNORE_ID = -1  # openeat/utils/common.py:24


def default_fbank_fn():
    try:
        import torch
        import torchaudio.compliance.kaldi as kaldi

        def fn(waveform, mel_bins, dither, sample_rate):
            mat = kaldi.fbank(torch.from_numpy(np.ascontiguousarray(waveform, dtype=np.float32))[None],
                              num_mel_bins=mel_bins, frame_length=25, frame_shift=10, dither=dither,
                              energy_floor=0.0, sample_frequency=sample_rate)
            return mat.detach().numpy()
        return fn, 'torchaudio.compliance.kaldi.fbank'
    except Exception:  # pragma: no cover - torchaudio missing
        def fn(waveform, mel_bins, dither, sample_rate):
            assert dither == 0.0, 'numpy oracle has no dither'
            return ofbank.fbank(waveform, num_mel_bins=mel_bins, sample_frequency=float(sample_rate))
        return fn, 'oracle.fbank.fbank (numpy)'


def default_speed_fn():
    """Speed perturb for the timed CPU baseline: ``torchaudio.functional.speed`` (conv1d-based, the
    substitute oracle itself) when torchaudio is importable, else the numpy restatement."""
    try:
        import torch
        import torchaudio.functional as TF

        def fn(waveform, sample_rate, speed):
            if speed == 1.0:
                return waveform
            y, _ = TF.speed(torch.from_numpy(np.ascontiguousarray(waveform, dtype=np.float32))[None], sample_rate, speed)
            return y[0].numpy()
        return fn
    except Exception:  # pragma: no cover
        return ospeed.speed_perturb


def read_wav(path, start=None, end=None):
    """16-bit PCM mono/multi-channel -> (float32 int16-scale samples of channel 0, sample_rate).
    Mirrors dataset.py:62-75 (``start``/``end`` in seconds -> frame_offset/num_frames)."""
    with wave.open(path, 'rb') as w:
        sr = w.getframerate()
        assert w.getsampwidth() == 2, 'only 16-bit PCM'
        nch = w.getnchannels()
        if start is not None:
            s = int(float(start) * sr)
            e = int(float(end) * sr)
            w.setpos(min(s, w.getnframes()))
            raw = w.readframes(max(0, e - s))
        else:
            raw = w.readframes(w.getnframes())
    pcm = np.frombuffer(raw, dtype='<i2')
    if nch > 1:
        pcm = pcm.reshape(-1, nch)[:, 0]
    return pcm.astype(np.float32), sr


def extract_feature(batch, conf, fbank_fn=None, rng=random, speed_fn=None):
    """dataset.py:39-118."""
    if fbank_fn is None:
        fbank_fn = default_fbank_fn()[0]
    if speed_fn is None:
        speed_fn = ospeed.speed_perturb
    speed_perturb_rate = conf.get('speed_perturb_rate', 0.5)
    speeds = conf.get('speeds', None)
    keys, feats, lengths, labels = [], [], [], []
    for x in batch:
        try:
            wav = x[1]
            if isinstance(wav, str):
                value = wav.strip().split(",")
                assert len(value) == 1 or len(value) == 3
                if len(value) == 3:
                    waveform, sample_rate = read_wav(value[0], value[1], value[2])
                else:
                    waveform, sample_rate = read_wav(value[0])
            else:                       # in-memory (samples, sample_rate)
                waveform, sample_rate = np.asarray(wav[0], dtype=np.float32), wav[1]
            resample_rate = conf.get('resample_rate', sample_rate)
            if resample_rate != sample_rate:
                g = np.gcd(int(sample_rate), int(resample_rate))
                waveform = ospeed.resample(waveform, int(sample_rate) // g, int(resample_rate) // g)
                sample_rate = resample_rate
            speed = x[3]
            if rng.random() < speed_perturb_rate:
                speed = ospeed.speed_generator(speeds, rng)
            if speed != 1.0:
                waveform = speed_fn(waveform, sample_rate, speed)
            mat = fbank_fn(waveform, conf['mel_bins'], conf['wav_dither'], sample_rate)
            feats.append(mat)
            keys.append(x[0])
            lengths.append(mat.shape[0])
            labels.append(np.array(x[2]))
        except Exception as e:  # dataset.py:108-111
            print(e)
            logging.warning('read utterance {} error'.format(x[0]))
    order = np.argsort(lengths)[::-1]
    return [keys[i] for i in order], [feats[i] for i in order], [labels[i] for i in order]


def pad_list(arrs, pad_value, dtype):
    """torch pad_sequence(batch_first=True) for 1-D / 2-D arrays."""
    n = max(a.shape[0] for a in arrs)
    out = np.full((len(arrs), n) + arrs[0].shape[1:], pad_value, dtype=dtype)
    for i, a in enumerate(arrs):
        out[i, :a.shape[0]] = a
    return out


class AudioCollate(object):
    """dataset.py:155-239 for data_type='wav' (numpy outputs instead of torch tensors)."""

    def __init__(self, feature_dither=0.0, spec_aug=False, spec_aug_conf=None, spec_sub=False,
                 spec_sub_conf=None, data_type='wav', feature_extraction_conf=None, normalization=True,
                 fbank_fn=None, rng=random, speed_fn=None):
        assert data_type == 'wav'
        assert feature_dither == 0.0, 'feature dither is stochastic: no parity claim (SURVEY 8a a6)'
        self.spec_aug, self.spec_aug_conf = spec_aug, spec_aug_conf or {}
        self.spec_sub, self.spec_sub_conf = spec_sub, spec_sub_conf or {}
        self.conf = feature_extraction_conf
        self.normalization = normalization
        self.fbank_fn = fbank_fn
        self.speed_fn = speed_fn
        self.rng = rng

    def __call__(self, batch):
        if len(batch) == 1:
            batch = batch[0]
        keys, xs, ys = extract_feature(batch, self.conf, self.fbank_fn, self.rng, self.speed_fn)
        if self.normalization:
            xs = [augment.normalization(x) for x in xs]
        if self.spec_sub:
            xs = [augment.spec_substitute(x, rng=self.rng, **self.spec_sub_conf) for x in xs]
        if self.spec_aug:
            xs = [augment.spec_augmentation(x, rng=self.rng, **self.spec_aug_conf) for x in xs]
        features_length = np.array([x.shape[0] for x in xs], dtype=np.int32)
        if len(xs) > 0:
            features = pad_list([np.asarray(x, dtype=np.float32) for x in xs], 0, np.float32)
            targets = pad_list([np.asarray(y, dtype=np.int32) for y in ys], IGNORE_ID, np.int32)
        else:
            features = np.zeros((0,), np.float32)
            targets = np.zeros((0,), np.float32)
        targets_length = np.array([y.shape[0] for y in ys], dtype=np.int32)
        return keys, {'features': features, 'features_length': features_length,
                      'targets': targets, 'targets_length': targets_length}
