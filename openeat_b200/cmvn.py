"""Global CMVN: drop-ins for ``openeat/modules/cmvn.py`` (GlobalCMVN) and ``openeat/utils/cmvn.py``
(load_cmvn), plus ``compute_cmvn_stats`` -- absent in the reference, whose stats-file format
(``openeat/utils/cmvn.py:30-35``) is the contract this writer honours.
"""
import json
import math

import numpy as np
import torch

from .frontend import default_frontend


def _stats_to_cmvn(means, variance, count):
    """openeat/utils/cmvn.py:36-43 / 78-85 (Python float64 arithmetic)."""
    for i in range(len(means)):
        means[i] /= count
        variance[i] = variance[i] / count - means[i] * means[i]
        if variance[i] < 1.0e-20:
            variance[i] = 1.0e-20
        variance[i] = 1.0 / math.sqrt(variance[i])
    return np.array([means, variance])


def _load_json_cmvn(json_cmvn_file):
    """openeat/utils/cmvn.py:21-43."""
    with open(json_cmvn_file) as f:
        cmvn_stats = json.load(f)
    return _stats_to_cmvn(cmvn_stats['mean_stat'], cmvn_stats['var_stat'], cmvn_stats['frame_num'])


def _load_kaldi_cmvn(kaldi_cmvn_file):
    """openeat/utils/cmvn.py:46-85 (text format; a binary file raises instead of sys.exit)."""
    with open(kaldi_cmvn_file, 'r') as fid:
        if fid.read(2) == '\0B':
            raise ValueError('kaldi cmvn binary file is not supported, please recompute it by: '
                             'compute-cmvn-stats --binary=false scp:feats.scp global_cmvn')
        fid.seek(0)
        arr = fid.read().split()
    assert arr[0] == '['
    assert arr[-2] == '0'
    assert arr[-1] == ']'
    feat_dim = int((len(arr) - 2 - 2) / 2)
    means = [float(arr[i]) for i in range(1, feat_dim + 1)]
    count = float(arr[feat_dim + 1])
    variance = [float(arr[i]) for i in range(feat_dim + 2, 2 * feat_dim + 2)]
    return _stats_to_cmvn(means, variance, count)


def load_cmvn(cmvn_file, is_json):
    """openeat/utils/cmvn.py:88-93 -> (mean, istd) float64 arrays."""
    cmvn = _load_json_cmvn(cmvn_file) if is_json else _load_kaldi_cmvn(cmvn_file)
    return cmvn[0], cmvn[1]


class GlobalCMVN(torch.nn.Module):
    """openeat/modules/cmvn.py:18-46.  Same constructor, same ``mean`` / ``istd`` buffers (so OpenEAT /
    wenet checkpoints load, openeat/utils/checkpoint.py:19-21); ``forward`` is the ``oe_cmvn_apply``
    CUDA kernel.  Inputs must be fp32 CUDA tensors: there is no CPU path."""

    def __init__(self, mean: torch.Tensor, istd: torch.Tensor, norm_var: bool = True):
        super().__init__()
        assert mean.shape == istd.shape
        self.norm_var = norm_var
        self.register_buffer("mean", mean)
        self.register_buffer("istd", istd)

    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError('openeat_b200.GlobalCMVN runs on CUDA tensors only (no CPU fallback)')
        if x.dtype != torch.float32 or self.mean.dtype != torch.float32:
            raise RuntimeError('openeat_b200.GlobalCMVN expects fp32 features and buffers')
        fe = default_frontend(mel_bins=x.shape[-1], device=x.device)
        return fe.cmvn_apply(x, self.mean.to(x.device), self.istd.to(x.device) if self.norm_var else None)


def write_json_cmvn(path, mean_stat, var_stat, frame_num):
    """Writes the JSON stats format parsed at openeat/utils/cmvn.py:30-35."""
    with open(path, 'w') as f:
        json.dump({'mean_stat': [float(v) for v in mean_stat], 'var_stat': [float(v) for v in var_stat],
                   'frame_num': int(frame_num)}, f)


def all_reduce_stats(stats, group=None):
    """The one collective of the path: sum of (sum[F], sumsq[F], count) = 2F+1 doubles over the ranks
    (NCCL over NVLink on GPUs, gloo in the CPU tests).  No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def compute_cmvn_stats(batches, mel_bins=80, sample_rate=16000, frontend=None, group=None, out_json=None):
    """Sum, sum of squares and frame count of the raw log-mel frames (dither 0, no normalisation, no
    augmentation) over this rank's utterances, all-reduced over the process group.

    batches: iterable of lists of int16 / fp32 waveforms (numpy, int16 scale) -- this rank's shard.
    Returns (mean_stat[F], var_stat[F], frame_num) as float64 numpy / int; rank 0 (or a single
    process) writes ``out_json`` in the format ``load_cmvn(..., is_json=True)`` reads.
    """
    import torch.distributed as dist

    from .frontend import pack_waveforms
    fe = frontend or default_frontend(mel_bins, sample_rate)
    F = fe.mel_bins
    stats = torch.zeros(2 * F + 1, dtype=torch.float64, device=fe.device)
    for waves in batches:
        if not len(waves):
            continue
        dtype = np.float32 if np.asarray(waves[0]).dtype.kind == 'f' else np.int16
        buf, offs, lens = pack_waveforms(waves, dtype=dtype)
        fe.fbank(buf.to(fe.device, non_blocking=True), offs, lens, layout='ragged', stats=stats, want_out=False)
    all_reduce_stats(stats, group)
    host = stats.cpu().numpy()
    mean_stat, var_stat, frame_num = host[:F], host[F:2 * F], int(round(host[2 * F]))
    rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank(group) == 0
    if out_json and rank0:
        write_json_cmvn(out_json, mean_stat, var_stat, frame_num)
    return mean_stat, var_stat, frame_num
