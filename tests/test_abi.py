"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol ``include/openeat_frontend.h`` declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from openeat_b200 import _lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'openeat_frontend.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(oe_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    _lib.build()
    return _lib.load()


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS), 'ctypes table and header drifted apart'


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_host_only_entry_points(lib):
    assert lib.oe_abi_version() == 3
    cfg = _lib.OeConfig()
    assert lib.oe_config_default(ctypes.byref(cfg)) == 0
    assert (cfg.sample_rate, cfg.frame_length, cfg.frame_shift, cfg.fft_size, cfg.num_mel_bins) == (16000, 400, 160, 512, 80)
    assert abs(cfg.preemph - 0.97) < 1e-7 and abs(cfg.log_floor - 1.1920928955078125e-07) < 1e-12
    # kaldi.py:63-67 frame count and torchaudio functional.py:1427 output length
    assert [lib.oe_num_frames(None, n) for n in (0, 399, 400, 559, 560, 80000, 560000)] == [0, 0, 1, 1, 2, 498, 3498]
    assert lib.oe_resample_out_len(80000, 9, 10) == 88889 and lib.oe_resample_out_len(80000, 11, 10) == 72728
    assert lib.oe_config_default(None) != 0 and b'null' in lib.oe_last_error()


def test_struct_layout_matches_header():
    """Field order of the ctypes mirrors == field order in the header."""
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    for cname, cls in (('oe_config', _lib.OeConfig), ('oe_batch', _lib.OeBatch), ('oe_resample_batch', _lib.OeResampleBatch)):
        body = re.search(r'typedef struct \{([^{}]*)\}\s*' + cname + ';', src).group(1)
        fields = [re.search(r'(\w+)\s*$', d.strip()).group(1) for d in body.split(';') if d.strip()]
        assert fields == [f[0] for f in cls._fields_], cname
    body = re.search(r'typedef struct oe_flac_frame \{([^{}]*)\}\s*oe_flac_frame;', src).group(1)
    fields = [n.strip() for d in body.split(';') if d.strip() for n in re.sub(r'^\s*\w+\s+', '', d.strip()).split(',')]
    assert fields == [f[0] for f in _lib.OeFlacFrame._fields_]
    from openeat_b200.ingest import FRAME_BYTES
    assert ctypes.sizeof(_lib.OeFlacFrame) == FRAME_BYTES == 48


def test_product_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from openeat_b200 import FrontendError
    from openeat_b200.frontend import Frontend
    with pytest.raises(FrontendError):
        Frontend()
