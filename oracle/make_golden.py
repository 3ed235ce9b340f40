"""Generates the committed fixtures in ``tests/golden/`` (oracle tooling).

Run in the BUILD container only (needs ``/root/reference`` and torchaudio):

    python -m oracle.make_golden

Sources of truth, in order of authority:
* the reference's own functions, imported unmodified under ``oracle/ref_shim.py``
  (``feature_processor``, ``utils.cmvn``, ``modules.cmvn``, ``dataset``);
* ``torchaudio.compliance.kaldi.fbank`` 2.11.0 -- the third-party function the
  reference calls at ``openeat/dataset/dataset.py:93-100`` -- in fp32 and fp64;
* ``torchaudio.functional.speed`` as the substitute speed-perturb oracle.
Nothing here is read at test time on the GPU box; only the ``.npz``/``.json``
outputs are.
"""
import json
import os
import random
import tempfile
import wave

import numpy as np

from . import ref_shim, signals

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

FBANK_CASES = [  # (class, samples, seed)
    ('white', 400, 11), ('white', 559, 12), ('white', 560, 13), ('white', 8000, 14),
    ('white', 80000, 15), ('white', 560000, 16),
    ('speech', 8000, 21), ('speech', 80000, 22), ('lsb', 8000, 31), ('zero', 8000, 0),
    ('dcsine', 8000, 0), ('square', 8000, 0),
]


def write_wav(path, pcm, sr=16000):
    with wave.open(path, 'wb') as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.asarray(pcm, dtype='<i2').tobytes())


def main():
    import torch
    import torchaudio.compliance.kaldi as kaldi
    import torchaudio.functional as TF

    ref_shim.install()
    from openeat.dataset import feature_processor as ref_fp
    from openeat.utils import cmvn as ref_cmvn
    from openeat.modules.cmvn import GlobalCMVN
    from openeat.dataset import dataset as ref_ds
    from openeat.dataset.audio_processor import _speed_generator

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)

    # 1. tables exactly as torchaudio builds them (kaldi.py:98-100, 436-511)
    window = kaldi._feature_window_function('povey', 400, 0.42, torch.device('cpu'), torch.float32).numpy()
    mel = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)[0].numpy()
    np.savez_compressed(os.path.join(OUT, 'tables.npz'), window=window, mel=mel)

    # 2. fbank: fp32 and fp64 torchaudio outputs on every signal class / edge length
    def tfbank(x, dt):
        return kaldi.fbank(torch.from_numpy(x.astype(np.float64)).to(dt)[None], num_mel_bins=80,
                           frame_length=25, frame_shift=10, dither=0.0, energy_floor=0.0,
                           sample_frequency=16000).numpy()

    manifest = {}
    for kind, n, seed in FBANK_CASES:
        x = signals.make(kind, n, seed)
        y32, y64 = tfbank(x, torch.float32), tfbank(x, torch.float64)
        name = 'fbank_%s_%d' % (kind, n)
        stride = 16 if n > 100000 else 1          # keep the long case small: every 16th frame
        np.savez_compressed(os.path.join(OUT, name + '.npz'),
                            y32=y32[::stride], y64=y64[::stride].astype(np.float32),
                            colsum64=y64.sum(axis=0), stride=stride, frames=y32.shape[0])
        manifest[name] = {'kind': kind, 'samples': n, 'seed': seed, 'sha256': signals.digest(x),
                          'frames': int(y32.shape[0]), 'stride': stride,
                          'gap32_64': float(np.abs(y32 - y64).max())}
    # the reference drops utterances shorter than one window (kaldi.py:142 raises)
    try:
        tfbank(signals.make('white', 399, 10), torch.float32)
        manifest['short_399_raises'] = False
    except AssertionError:
        manifest['short_399_raises'] = True

    # 3. reference augmentation functions with seeded Python RNG
    aug = {}
    rng = np.random.default_rng(77)
    for i, (T, F) in enumerate([(498, 80), (48, 80), (7, 80), (1, 80), (298, 40)]):
        x = rng.normal(10.0, 3.0, (T, F)).astype(np.float32)
        aug['x%d' % i] = x
        random.seed(1000 + i)
        aug['aug%d' % i] = ref_fp._spec_augmentation(x, num_t_mask=3, num_f_mask=2, max_t=50, max_f=10)
        random.seed(2000 + i)
        aug['sub%d' % i] = ref_fp._spec_substitute(x, max_t=30, num_t_sub=3)
        random.seed(3000 + i)                    # chained exactly like dataset.py:204-209
        aug['subaug%d' % i] = ref_fp._spec_augmentation(ref_fp._spec_substitute(x, max_t=30, num_t_sub=3),
                                                        num_t_mask=3, num_f_mask=2, max_t=50, max_f=10)
        if T > 1:
            aug['norm%d' % i] = ref_fp._normalization(x)
    random.seed(5)
    aug['speed_draws'] = np.array([_speed_generator([0.9, 1.1, 0.1]) for _ in range(8)]
                                  + [_speed_generator(None)] + [_speed_generator([1.05])])
    random.seed(6)
    aug['speed_draws_uniform'] = np.array([_speed_generator([0.9, 1.1, 0]) for _ in range(8)])
    np.savez_compressed(os.path.join(OUT, 'augment.npz'), **aug)

    # 4. CMVN: stats files -> reference load_cmvn -> GlobalCMVN
    feats = [tfbank(signals.make('speech', n, s), torch.float32) for n, s in [(8000, 41), (12000, 42), (5000, 43)]]
    s64 = sum(f.astype(np.float64).sum(0) for f in feats)
    q64 = sum((f.astype(np.float64) ** 2).sum(0) for f in feats)
    cnt = sum(f.shape[0] for f in feats)
    with open(os.path.join(OUT, 'cmvn_stats.json'), 'w') as f:
        json.dump({'mean_stat': s64.tolist(), 'var_stat': q64.tolist(), 'frame_num': cnt}, f)
    with open(os.path.join(OUT, 'cmvn_stats.kaldi.txt'), 'w') as f:
        f.write('[ ' + ' '.join(repr(float(v)) for v in s64) + ' ' + repr(float(cnt)) + '\n')
        f.write(' '.join(repr(float(v)) for v in q64) + ' 0 ]\n')
    mean_j, istd_j = ref_cmvn.load_cmvn(os.path.join(OUT, 'cmvn_stats.json'), True)
    mean_k, istd_k = ref_cmvn.load_cmvn(os.path.join(OUT, 'cmvn_stats.kaldi.txt'), False)
    gc = GlobalCMVN(torch.from_numpy(mean_j).float(), torch.from_numpy(istd_j).float())
    xb = np.zeros((2, feats[1].shape[0], 80), np.float32)          # padded batch, pads are 0
    xb[0, :feats[0].shape[0]] = feats[0]
    xb[1] = feats[1]
    np.savez_compressed(os.path.join(OUT, 'cmvn.npz'), mean_json=mean_j, istd_json=istd_j, mean_kaldi=mean_k,
                        istd_kaldi=istd_k, x=xb, y=gc(torch.from_numpy(xb)).numpy(),
                        y_novar=GlobalCMVN(torch.from_numpy(mean_j).float(), torch.from_numpy(istd_j).float(),
                                           norm_var=False)(torch.from_numpy(xb)).numpy(),
                        sum=s64, sumsq=q64, count=cnt)

    # 5. the reference's own audio_collate_func, end to end, under the shim
    tmp = tempfile.mkdtemp()
    lens = [16000, 9000, 5200, 12345, 300, 7777]          # 300 < 400 -> dropped by the reference
    pcm = [signals.make('speech' if i % 2 else 'white', n, 50 + i) for i, n in enumerate(lens)]
    batch = []
    for i, p in enumerate(pcm):
        path = os.path.join(tmp, 'u%d.wav' % i)
        write_wav(path, p)
        batch.append(('utt%d' % i, path, [i + 1] * (i + 2), 1.0))
    batch.append(('seg', os.path.join(tmp, 'u0.wav') + ',0.25,0.75', [9, 9], 1.0))   # segmented entry
    conf = {'resample_rate': 16000, 'speed_perturb_rate': 0, 'speeds': [0.9, 1.1, 0.1], 'wav_dither': 0.0,
            'mel_bins': 80}
    col = {'pcm%d' % i: p for i, p in enumerate(pcm)}
    for tag, kw in [('plain', dict(normalization=False)),
                    ('norm_aug', dict(normalization=True, spec_aug=True,
                                      spec_aug_conf=dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10))),
                    ('sub_aug', dict(normalization=False, spec_sub=True, spec_sub_conf=dict(num_t_sub=3, max_t=30),
                                     spec_aug=True, spec_aug_conf=dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10)))]:
        fn = ref_ds.audio_collate_func(data_type='wav', feature_extraction_conf=conf, **kw)
        random.seed(4242)
        keys, out = fn([batch])
        col[tag + '_keys'] = np.array(keys)
        for k2, v in out.items():
            col[tag + '_' + k2] = v.numpy()
    np.savez_compressed(os.path.join(OUT, 'collate.npz'), **col)

    # 6. substitute speed oracle: torchaudio.functional.speed
    x = signals.make('speech', 8000, 61).astype(np.float32)
    sp = {'x': x}
    for s, tag in [(0.9, '090'), (1.1, '110')]:
        sp['y' + tag] = TF.speed(torch.from_numpy(x)[None], 16000, s)[0][0].numpy()
        sp['fb' + tag] = tfbank(sp['y' + tag], torch.float32)
    np.savez_compressed(os.path.join(OUT, 'speed.npz'), **sp)

    # 7. the reference's AudioDataset (list parsing + batching) on a synthetic format.data
    import zhon.hanzi                                # shim module: give it a punctuation set so the regex compiles
    zhon.hanzi.punctuation = '\u3002\uff0c'
    rng = np.random.default_rng(88)
    words = ['HELLO', 'WORLD', '\u4f60', '\u597d', 'OKAY', '\u7684', 'A', 'B']
    char_dict = {w: i for i, w in enumerate(['<blank>', '<unk>'] + words + ['<sos/eos>'])}
    lines = []
    for i in range(40):
        secs = float(np.round(rng.uniform(0.05, 21.0), 2))
        nw = int(rng.integers(1, 6))
        text = ' '.join(words[j] for j in rng.integers(0, len(words), nw))
        lines.append('utt:u%03d\tfeat:/data/wav/u%03d.wav\tfeat_shape:%s\ttext:%s' % (i, i, secs, text))
    lines.insert(5, 'utt:seg0\tfeat:/data/wav/long.wav,1.5,4.25\tfeat_shape:2.75\ttext:HELLO \u4f60')
    lines.insert(9, 'malformed line without tabs')
    list_path = os.path.join(OUT, 'format.data')
    with open(list_path, 'w', encoding='utf-8') as f:
        f.write('\n'.join(lines) + '\n')
    with open(os.path.join(OUT, 'format.dict.json'), 'w', encoding='utf-8') as f:
        json.dump(char_dict, f, ensure_ascii=False)
    ds_gold = {}
    for tag, kw in [('static', dict(batch_type='static', batch_size=4, sort=True, max_length=2000, min_length=10)),
                    ('dynamic', dict(batch_type='dynamic', max_frames_in_batch=3000, sort=True, max_length=2000, min_length=10)),
                    ('dynamic_unsorted_speed', dict(batch_type='dynamic', max_frames_in_batch=1500, sort=False,
                                                    speed_perturb=True, max_length=2000, min_length=10)),
                    ('shuffle', dict(batch_type='shuffle', batch_size=8, sort=False))]:
        ds = ref_ds.AudioDataset(list_path, char_dict, None, data_type='wav', **kw)
        ds_gold[tag] = [[[x[0], x[1], list(x[2]), float(x[3])] for x in (b if tag != 'shuffle' else [b])] for b in ds.data]
        ds_gold[tag + '_batch_size'] = ds.batch_size
    with open(os.path.join(OUT, 'audio_dataset.json'), 'w', encoding='utf-8') as f:
        json.dump(ds_gold, f, ensure_ascii=False)

    # 8. the reference's audio_collate_func on Kaldi-archive features (data_type != 'wav': _load_feature,
    #    dataset.py:120-152, then the common chain :195-238), under the shim.  kaldi_io is a third-party package
    #    that is not installed here: the shim's stub module gets read_mat from openeat_b200.kaldi_io, so this pins
    #    the reference's handling of the matrices (sort order, the doubled label list, normalisation, masks,
    #    padding) -- the binary format itself is only pinned by Kaldi's published layout (no Kaldi binaries here).
    import sys
    from openeat_b200 import kaldi_io as my_kaldi_io
    sys.modules['kaldi_io'].read_mat = my_kaldi_io.read_mat
    rng = np.random.default_rng(99)
    T = [57, 120, 33, 240, 5, 120]
    mats = [rng.normal(3.0, 2.0, (t, 80)).astype(np.float32) for t in T]
    scp = my_kaldi_io.write_mat_ark(os.path.join(tmp, 'feats.ark'), [('k%d' % i, m) for i, m in enumerate(mats)])
    kbatch = [('k%d' % i, scp['k%d' % i], [i + 1] * (i % 3 + 1), 1.0) for i in range(len(T))]
    kc = {'mat%d' % i: m for i, m in enumerate(mats)}
    kc['labels'] = np.array([len(x[2]) for x in kbatch])
    for tag, kw in [('plain', dict(normalization=False)),
                    ('norm_aug', dict(normalization=True, spec_aug=True,
                                      spec_aug_conf=dict(num_t_mask=2, num_f_mask=2, max_t=20, max_f=8))),
                    ('sub', dict(normalization=True, spec_sub=True, spec_sub_conf=dict(num_t_sub=3, max_t=10)))]:
        fn = ref_ds.audio_collate_func(data_type='kaldi', **kw)
        random.seed(777)
        keys, out = fn([kbatch])
        kc[tag + '_keys'] = np.array(keys)
        for k2, v in out.items():
            kc[tag + '_' + k2] = v.numpy()
    np.savez_compressed(os.path.join(OUT, 'kaldi_collate.npz'), **kc)

    with open(os.path.join(OUT, 'manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print('wrote', sorted(os.listdir(OUT)))


if __name__ == '__main__':
    main()
