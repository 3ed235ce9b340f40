"""Round-2 additions to the committed fixtures in ``tests/golden/`` (oracle tooling; run in the BUILD container only):

    python -m oracle.make_golden_r02

* ``fbank_mel{23,40}_*.npz``: ``torchaudio.compliance.kaldi.fbank`` 2.11.0 (fp32 and fp64) with other ``num_mel_bins``
  (``feature_extraction_conf['mel_bins']``, dataset.py:95) plus torchaudio's own mel matrix for that bin count -- pins the
  table-driven first-generation kernel.
* ``resample_long.npz``: ``torchaudio.functional.resample`` (the function behind ``transforms.Resample``, dataset.py:77-84,
  and behind the substitute speed oracle) for ratios whose polyphase table is too long to keep: 44.1 kHz -> 16 kHz
  (441:160) and a speed drawn from a continuous range (953:1000) -- pins the on-the-fly (OE_RS_DIRECT) resampler.
"""
import os

import numpy as np

from . import signals

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    import torch
    import torchaudio.compliance.kaldi as kaldi
    import torchaudio.functional as TF
    torch.set_num_threads(1)
    for bins, kind, n, seed in ((23, 'speech', 8000, 51), (23, 'white', 560, 52), (40, 'white', 8000, 53), (40, 'dcsine', 8000, 0)):
        x = signals.make(kind, n, seed)

        def run(dt):
            return kaldi.fbank(torch.from_numpy(x.astype(np.float64)).to(dt)[None], num_mel_bins=bins, frame_length=25,
                               frame_shift=10, dither=0.0, energy_floor=0.0, sample_frequency=16000).numpy()
        mel = kaldi.get_mel_banks(bins, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)[0].numpy()
        np.savez_compressed(os.path.join(OUT, 'fbank_mel%d_%s_%d.npz' % (bins, kind, n)), pcm=x, y32=run(torch.float32),
                            y64=run(torch.float64).astype(np.float32), mel=mel)
    out = {}
    x = signals.make('speech', 44100, 61).astype(np.float32)
    out['x_441_160'] = x
    out['y_441_160'] = TF.resample(torch.from_numpy(x)[None], 44100, 16000)[0].numpy()
    x = signals.make('speech', 8000, 62).astype(np.float32)
    out['x_953_1000'] = x
    out['y_953_1000'] = TF.resample(torch.from_numpy(x)[None], 953, 1000)[0].numpy()      # speed 0.953 at 16 kHz: 15248:16000
    np.savez_compressed(os.path.join(OUT, 'resample_long.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
