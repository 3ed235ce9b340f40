"""Numpy restatement of the global-CMVN contract (oracle; tests only).

* ``load_cmvn``     -- ``openeat/utils/cmvn.py:21-93`` (JSON and Kaldi-text stats).
* ``global_cmvn``   -- ``openeat/modules/cmvn.py:35-46`` (two separate fp32 ops).
* ``compute_cmvn_stats`` / ``write_json_cmvn`` -- ABSENT in the reference; the file
  format read by ``_load_json_cmvn`` (cmvn.py:30-35) is the contract, so the stats
  are the sum, sum of squares and count of the raw (un-normalised, un-augmented,
  dither-0) log-mel frames.
"""
import json
import math

import numpy as np


def stats_to_mean_istd(mean_stat, var_stat, frame_num):
    """cmvn.py:36-41 / 78-83 in Python float64.  Returns float64 arrays."""
    means = [float(v) for v in mean_stat]
    variance = [float(v) for v in var_stat]
    count = frame_num
    for i in range(len(means)):
        means[i] /= count
        variance[i] = variance[i] / count - means[i] * means[i]
        if variance[i] < 1.0e-20:
            variance[i] = 1.0e-20
        variance[i] = 1.0 / math.sqrt(variance[i])
    return np.array(means), np.array(variance)


def load_json_cmvn(path):
    """cmvn.py:21-43."""
    with open(path) as f:
        st = json.load(f)
    return stats_to_mean_istd(st['mean_stat'], st['var_stat'], st['frame_num'])


def load_kaldi_cmvn(path):
    """cmvn.py:46-85 (text format ``[ sums... count \\n sumsq... 0 ]``)."""
    with open(path, 'r') as fid:
        if fid.read(2) == '\0B':
            raise ValueError('kaldi cmvn binary file is not supported')
        fid.seek(0)
        arr = fid.read().split()
    assert arr[0] == '[' and arr[-2] == '0' and arr[-1] == ']'
    feat_dim = int((len(arr) - 2 - 2) / 2)
    means = [float(arr[i]) for i in range(1, feat_dim + 1)]
    count = float(arr[feat_dim + 1])
    variance = [float(arr[i]) for i in range(feat_dim + 2, 2 * feat_dim + 2)]
    return stats_to_mean_istd(means, variance, count)


def load_cmvn(path, is_json):
    """cmvn.py:88-93."""
    return load_json_cmvn(path) if is_json else load_kaldi_cmvn(path)


def global_cmvn(x, mean, istd, norm_var=True):
    """modules/cmvn.py:43-46 on fp32 (mean/istd cast to fp32 at asr_model.py:82-83)."""
    x = np.asarray(x, dtype=np.float32)
    y = x - np.asarray(mean, dtype=np.float32)
    if norm_var:
        y = y * np.asarray(istd, dtype=np.float32)
    return y


def compute_cmvn_stats(feature_list):
    """Sum, sum of squares (float64) and frame count over a list of (T_i, F) arrays."""
    dim = feature_list[0].shape[1]
    s = np.zeros(dim, dtype=np.float64)
    q = np.zeros(dim, dtype=np.float64)
    n = 0
    for f in feature_list:
        f64 = np.asarray(f, dtype=np.float64)
        s += f64.sum(axis=0)
        q += (f64 * f64).sum(axis=0)
        n += f.shape[0]
    return s, q, n


def write_json_cmvn(path, mean_stat, var_stat, frame_num):
    """Writer for the JSON stats format parsed at cmvn.py:30-35."""
    with open(path, 'w') as f:
        json.dump({'mean_stat': [float(v) for v in mean_stat],
                   'var_stat': [float(v) for v in var_stat],
                   'frame_num': int(frame_num)}, f)
