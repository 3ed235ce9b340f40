"""Import shim that lets the UNMODIFIED reference modules run in the build container
(oracle tooling; used only by ``oracle/make_golden.py``; never on the GPU box).

``/root/reference`` targets torch 1.9 / torchaudio 0.9 + libsox + kaldi_io + zhon.
Here (torchaudio 2.11, no libsox, no TorchCodec) the shim supplies:

* ``torchaudio.set_audio_backend``        -> no-op  (audio_processor.py:4, dataset.py:28)
* ``kaldi_io`` / ``zhon.hanzi``           -> stub modules (dataset.py:21, text_processor.py:25)
* ``torchaudio.backend.sox_io_backend.info`` and ``torchaudio.load`` -> stdlib
  ``wave`` readers returning ``int16 / 32768`` as fp32, exactly what libsox gives
  (dataset.py:62-72)

Nothing in the reference is edited or copied; the shim only patches the process.
"""
import sys
import types
import wave

import numpy as np

REFERENCE_ROOT = '/root/reference'


def install():
    import torch
    import torchaudio

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    torchaudio.set_audio_backend = lambda *_a, **_k: None
    sys.modules.setdefault('kaldi_io', types.ModuleType('kaldi_io'))
    zhon = types.ModuleType('zhon')
    hanzi = types.ModuleType('zhon.hanzi')
    hanzi.punctuation = ''
    zhon.hanzi = hanzi
    sys.modules.setdefault('zhon', zhon)
    sys.modules.setdefault('zhon.hanzi', hanzi)

    class _Info(object):
        def __init__(self, sr, n):
            self.sample_rate, self.num_frames = sr, n

    def info(path):
        with wave.open(path, 'rb') as w:
            return _Info(w.getframerate(), w.getnframes())

    def load(filepath, num_frames=-1, frame_offset=0):
        with wave.open(filepath, 'rb') as w:
            sr, nch = w.getframerate(), w.getnchannels()
            w.setpos(min(frame_offset, w.getnframes()))
            n = w.getnframes() - frame_offset if num_frames < 0 else num_frames
            raw = w.readframes(max(0, n))
        pcm = np.frombuffer(raw, dtype='<i2').reshape(-1, nch).T
        return torch.from_numpy(pcm.astype(np.float32) / 32768.0), sr

    backend = types.ModuleType('torchaudio.backend')
    sox = types.ModuleType('torchaudio.backend.sox_io_backend')
    sox.info = info
    backend.sox_io_backend = sox
    torchaudio.backend = backend
    sys.modules['torchaudio.backend'] = backend
    sys.modules['torchaudio.backend.sox_io_backend'] = sox
    torchaudio.load = load
