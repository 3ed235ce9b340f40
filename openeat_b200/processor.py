"""wenet-UIO style processor chain over the CUDA front-end.

OpenEAT itself has no ``processor.py`` -- its front-end lives inline in ``dataset.py`` (mirrored by
``openeat_b200.dataset``).  The recipes it descends from (wenet) expose the same steps as generator
processors over sample dicts; BASELINE.json's north_star uses that vocabulary (``speed_perturb``,
``compute_fbank``, ``spec_aug``, ``spec_sub``, global CMVN, ``data.list``).  This module provides those
processor signatures.  Semantics follow OpenEAT's functions wherever both exist:

    speed_perturb     audio_processor.py:5-35 semantics per sample (``speeds`` = list to choose from, wenet
                      style; the draw is ``random.choice``), resampling on the GPU
    compute_fbank     dataset.py:93-100 (kaldi.fbank 25/10 ms, energy_floor 0) -- samples are grouped into
                      ragged GPU batches of ``batch_size`` internally, results are yielded one by one in order
    spec_sub/spec_aug feature_processor.py:10-64 (same random draws, Python ``random``)
    global_cmvn       modules/cmvn.py:35-46
    padding           dataset.py:211-231: sort by length (descending), zero-pad features, pad labels with -1

Sample dicts use wenet's keys: ``key``, ``wav`` (1-D int16 / float array or (1, N) tensor on the int16 scale /
[-1, 1) floats are NOT rescaled here), ``sample_rate``, ``label`` (token ids), ``feat`` ((T, F) CUDA tensor).
"""
import io
import json
import os
import random
import tarfile

import numpy as np
import torch

from .dataset import IGNORE_ID, decode_wav, read_wav
from .feature_processor import plan_spec_augmentation, plan_spec_substitute
from .frontend import default_frontend, pack_waveforms, speed_ratio


def parse_raw(lines):
    """data.list lines -> samples.  Accepts wenet ``data.list`` json lines ``{"key", "wav", "txt"}`` and
    OpenEAT ``format.data`` lines (``utt:<k>\\tfeat:<path[,start,end]>\\tfeat_shape:<sec>\\ttext:<t>``).
    Unreadable entries are skipped with a message, like dataset.py:108-111."""
    for line in lines:
        line = line.strip()
        if not line:
            continue
        try:
            if line.startswith('{'):
                obj = json.loads(line)
                key, wav, txt = obj['key'], obj['wav'], obj.get('txt', '')
                start, end = obj.get('start'), obj.get('end')
            else:
                arr = line.split('\t')
                key = arr[0].split(':')[1]
                wav = ':'.join(arr[1].split(':')[1:])
                txt = arr[3].split(':', 1)[1] if len(arr) > 3 else ''
                start = end = None
            value = wav.strip().split(',')
            if len(value) == 3:
                wav, start, end = value
            pcm, sr = read_wav(wav, start, end)
            yield {'key': key, 'wav': pcm, 'sample_rate': sr, 'txt': txt}
        except Exception as e:
            print(e)
            print('read utterance {} error'.format(line[:60]))


def tar_file_and_group(shards):
    """wenet shard lists: every line of ``data.list`` (``data_type='shard'``) names one tar file whose members are
    ``<key>.wav`` / ``<key>.flac`` (RIFF/WAVE or FLAC, see ``dataset.read_wav``) and ``<key>.txt``, stored next to each other.  Yields the same sample dicts as
    ``parse_raw``.  A member that cannot be decoded is skipped with a message (dataset.py:108-111 convention)."""
    for shard in shards:
        shard = shard.strip() if isinstance(shard, str) else shard
        if not shard:
            continue
        with tarfile.open(shard, 'r:*') as tar:
            cur, sample = None, {}
            for member in tar:
                if not member.isfile():
                    continue
                stem, _, ext = member.name.rpartition('.')
                if cur is not None and stem != cur:
                    if 'wav' in sample:
                        yield dict(sample, key=cur, txt=sample.get('txt', ''))
                    sample = {}
                cur = stem
                try:
                    blob = tar.extractfile(member).read()
                    if ext == 'txt':
                        sample['txt'] = blob.decode('utf8').strip()
                    elif ext in ('wav', 'flac'):
                        pcm, sr = decode_wav(io.BytesIO(blob), name=member.name)
                        sample['wav'] = np.array(pcm)
                        sample['sample_rate'] = sr
                except Exception as e:
                    print(e)
                    print('read utterance {} error'.format(member.name))
                    sample.pop('wav', None)
            if cur is not None and 'wav' in sample:
                yield dict(sample, key=cur, txt=sample.get('txt', ''))


def speed_perturb(data, speeds=None):
    """Per sample: pick a speed (``random.choice``) and resample on the GPU when it is not 1.0."""
    if speeds is None:
        speeds = [0.9, 1.0, 1.1]
    fe = default_frontend()
    for sample in data:
        speed = random.choice(speeds)
        if speed != 1.0:
            wav = np.asarray(sample['wav']).reshape(-1)
            buf, offs, lens = pack_waveforms([wav], dtype=np.float32 if wav.dtype.kind == 'f' else np.int16)
            out, ooffs, olens = fe.resample(buf.to(fe.device), offs, lens, [speed_ratio(speed, sample['sample_rate'])])
            sample = dict(sample, wav=out[ooffs[0]:ooffs[0] + int(olens[0])])
        yield sample


def compute_fbank(data, num_mel_bins=23, frame_length=25, frame_shift=10, dither=0.0, batch_size=64):
    """Log-mel filterbank of every sample (``feat``: (T, num_mel_bins) CUDA tensor).  Samples shorter than one
    window are dropped like the reference drops them (kaldi.py:142 raises, dataset.py:108-111)."""
    assert frame_length == 25 and frame_shift == 10, 'only 25 ms / 10 ms framing is built'
    fe = default_frontend(num_mel_bins)
    seed = [int.from_bytes(os.urandom(8), 'little')]         # wav dither: Philox key, advanced per GPU batch

    def flush(group):
        waves = []
        for s in group:
            w = s['wav']
            waves.append(w.detach().reshape(-1).cpu().numpy() if torch.is_tensor(w) else np.asarray(w).reshape(-1))
        any_f32 = any(w.dtype.kind == 'f' for w in waves)
        buf, offs, lens = pack_waveforms(waves, dtype=np.float32 if any_f32 else np.int16)
        seed[0] += 1
        out, frames = fe.fbank(buf.to(fe.device, non_blocking=True), offs, lens, layout='ragged',
                               wav_dither=float(dither), dither_seed=seed[0])
        r = 0
        for s, w, m in zip(group, waves, frames):
            if m == 0:
                print('choose a window size 400 that is [2, %d]' % len(w))
                continue
            s = dict(s, feat=out[r:r + int(m)])
            s.pop('wav', None)
            r += int(m)
            yield s

    group = []
    for sample in data:
        if sample.get('sample_rate', 16000) != 16000:
            print('sample rate %s is not supported by this front-end build' % sample.get('sample_rate'))
            continue
        group.append(sample)
        if len(group) == batch_size:
            for s in flush(group):
                yield s
            group = []
    if group:
        for s in flush(group):
            yield s


def _feat_ops(data, plan, batch_size=64):
    """Runs one post-fbank operation over groups of ``batch_size`` samples with ONE launch sequence per group (the
    features of a group are concatenated into a fresh, 16-byte aligned buffer; ``plan(frames, F)`` makes the group's
    random draws, sample by sample in order, and returns the keywords of ``Frontend.fbank(features_in=True)``)."""
    def flush(group):
        xs = [s['feat'] for s in group]
        F = int(xs[0].shape[1])
        fe = default_frontend(F)
        frames = np.array([x.shape[0] for x in xs], dtype=np.int32)
        kw = plan(frames, F)
        buf = torch.cat([x.to(fe.device, dtype=torch.float32) for x in xs]).contiguous()
        offs = np.concatenate([[0], np.cumsum(frames[:-1].astype(np.int64))]).astype(np.int64)
        out, _ = fe.fbank(buf, offs, frames, layout='ragged', features_in=True, **kw)
        r = 0
        for s, m in zip(group, frames):
            yield dict(s, feat=out[r:r + int(m)])
            r += int(m)

    group = []
    for sample in data:
        group.append(sample)
        if len(group) == batch_size:
            for s in flush(group):
                yield s
            group = []
    if group:
        for s in flush(group):
            yield s


def spec_sub(data, max_t=20, num_t_sub=3, batch_size=64):
    """feature_processor.py:44-64, one launch sequence per ``batch_size`` samples (draws: sample by sample, in order)."""
    return _feat_ops(data, lambda frames, F: dict(frame_maps=[plan_spec_substitute(int(t), max_t, num_t_sub) for t in frames]),
                     batch_size)


def spec_aug(data, num_t_mask=2, num_f_mask=2, max_t=50, max_f=10, batch_size=64):
    """feature_processor.py:10-42, one launch sequence per ``batch_size`` samples (draws: sample by sample, in order)."""
    def plan(frames, F):
        tm, fm = [], []
        for t in frames:
            a, b = plan_spec_augmentation(int(t), F, num_t_mask, num_f_mask, max_t, max_f)
            tm.append(a)
            fm.append(b)
        return dict(tmask=np.array(tm, np.int32).reshape(len(frames), -1, 2) if num_t_mask else None,
                    fmask=np.array(fm, np.int32).reshape(len(frames), -1, 2) if num_f_mask else None)
    return _feat_ops(data, plan, batch_size)


def utt_normalize(data, batch_size=64):
    """feature_processor.py:5-8 (OpenEAT's default ``normalization=True``), one launch sequence per ``batch_size`` samples."""
    return _feat_ops(data, lambda frames, F: dict(normalization=True), batch_size)


def global_cmvn(data, mean, istd, norm_var=True, batch_size=64):
    """modules/cmvn.py:35-46; ``mean`` / ``istd`` are fp32 tensors (any device); one launch sequence per ``batch_size`` samples."""
    dev = {}

    def plan(frames, F):
        if not dev:
            fe = default_frontend(F)
            dev['cmvn'] = (mean.to(fe.device, dtype=torch.float32).contiguous(),
                           istd.to(fe.device, dtype=torch.float32).contiguous() if norm_var else None)
        return dict(cmvn=dev['cmvn'])
    return _feat_ops(data, plan, batch_size)


def batch(data, batch_size=16):
    """Static batching: lists of ``batch_size`` samples."""
    buf = []
    for sample in data:
        buf.append(sample)
        if len(buf) >= batch_size:
            yield buf
            buf = []
    if buf:
        yield buf


def padding(data):
    """dataset.py:114-118, 211-231: per batch sort by feature length (descending, ``np.argsort(...)[::-1]``),
    zero-pad features, pad labels with -1.  Yields (keys, feats (B, Tmax, F), labels (B, Lmax), feat_lengths,
    label_lengths)."""
    for samples in data:
        lengths = [s['feat'].shape[0] for s in samples]
        order = np.argsort(lengths)[::-1]
        feats = [samples[i]['feat'] for i in order]
        keys = [samples[i]['key'] for i in order]
        labels = [torch.as_tensor(np.asarray(samples[i].get('label', []), dtype=np.int64)).int() for i in order]
        padded = torch.nn.utils.rnn.pad_sequence(feats, batch_first=True, padding_value=0)
        padded_labels = torch.nn.utils.rnn.pad_sequence(labels, batch_first=True, padding_value=IGNORE_ID)
        yield (keys, padded, padded_labels, torch.tensor([f.shape[0] for f in feats], dtype=torch.int32),
               torch.tensor([l.shape[0] for l in labels], dtype=torch.int32))
