"""Minimal Kaldi table I/O for the ``data_type='kaldi'`` path of the reference (``kaldi_io.read_mat`` at
openeat/dataset/dataset.py:138).  ``kaldi_io`` is a third-party package that is not vendored in the reference;
this restates the part the reference calls, from Kaldi's published binary matrix format:

    ark entry   = <key> ' ' '\\0' 'B' <matrix>
    <matrix>    = 'FM ' | 'DM '   then   '\\4' int32 rows  '\\4' int32 cols   then rows*cols little-endian float32 | float64
    rxspecifier = 'path/to/file.ark:byte_offset' -- the offset points at the '\\0B' header (what copy-feats writes
                  into feats.scp), or a plain path whose first entry is read

Compressed matrices ('CM', 'CM2', 'CM3') are not handled: ``read_mat`` raises ``ValueError`` for them.
"""
import struct

import numpy as np


def _read_matrix(fd):
    if fd.read(2) != b'\0B':
        raise ValueError('not a binary Kaldi object (text-mode archives are not supported)')
    tok = b''
    while True:                                   # token, e.g. b'FM ', terminated by a space
        ch = fd.read(1)
        if not ch:
            raise ValueError('unexpected end of file in the matrix header')
        tok += ch
        if ch == b' ':
            break
    if tok not in (b'FM ', b'DM '):
        raise ValueError('unsupported Kaldi matrix type %r' % tok)
    dims = []
    for _ in range(2):
        if fd.read(1) != b'\4':
            raise ValueError('bad dimension field')
        dims.append(struct.unpack('<i', fd.read(4))[0])
    rows, cols = dims
    dt = np.dtype('<f4') if tok == b'FM ' else np.dtype('<f8')
    buf = fd.read(rows * cols * dt.itemsize)
    if len(buf) != rows * cols * dt.itemsize:
        raise ValueError('truncated matrix')
    return np.frombuffer(buf, dtype=dt).reshape(rows, cols).copy()


def read_mat(rxfile):
    """``kaldi_io.read_mat``: one matrix from 'file.ark:offset' (or the first entry of a plain path) as a
    float32/float64 ndarray (rows, cols)."""
    path, offset = rxfile, None
    if ':' in rxfile and rxfile.rsplit(':', 1)[1].isdigit():
        path, off = rxfile.rsplit(':', 1)
        offset = int(off)
    with open(path, 'rb') as fd:
        if offset is not None:
            fd.seek(offset)
        else:                                     # skip the key of the first entry
            while True:
                ch = fd.read(1)
                if not ch:
                    raise ValueError('empty archive')
                if ch == b' ':
                    break
        return _read_matrix(fd)


def read_mat_ark(path):
    """Generator over (key, matrix) of a binary archive (``kaldi_io.read_mat_ark``)."""
    with open(path, 'rb') as fd:
        while True:
            key = b''
            while True:
                ch = fd.read(1)
                if not ch:
                    return
                if ch == b' ':
                    break
                key += ch
            yield key.decode(), _read_matrix(fd)


def write_mat_ark(path, items):
    """Writes (key, float32 matrix) pairs as a binary archive; returns {key: 'path:offset'} (the scp entries)."""
    scp = {}
    with open(path, 'wb') as fd:
        for key, mat in items:
            mat = np.ascontiguousarray(mat, dtype='<f4')
            fd.write(key.encode() + b' ')
            scp[key] = '%s:%d' % (path, fd.tell())
            fd.write(b'\0BFM ' + b'\4' + struct.pack('<i', mat.shape[0]) + b'\4' + struct.pack('<i', mat.shape[1]))
            fd.write(mat.tobytes())
    return scp
