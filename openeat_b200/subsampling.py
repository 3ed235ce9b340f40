"""The first consumer of the feature tensor, fused with GlobalCMVN (forward only).

``CmvnConvSubsample`` replaces the first two operations the reference's encoder applies to a batch
(``openeat/modules/encoder.py:221-223``): ``GlobalCMVN`` (``modules/cmvn.py:43-46``) and the first
``Conv2d(1, odim, 3, 2) + ReLU`` of ``Conv2dSubsampling4`` (``modules/subsampling.py:76-78, 110-111``).  The
normalised batch is never materialised: the CUDA kernel (``oe_cmvn_conv_subsample``) applies ``(x - mean) * istd`` while
it stages its input rows.  Everything behind that layer (second convolution, linear, positional encoding) is model code
and stays with the reference.
"""
import ctypes

import torch

from . import _lib
from ._lib import FrontendError, check


class CmvnConvSubsample(torch.nn.Module):
    """``forward(xs (B, T, F) fp32 CUDA) -> (B, odim, (T-3)//2+1, (F-3)//2+1)`` == ``relu(conv(global_cmvn(xs).unsqueeze(1)))``.

    ``conv``: the ``torch.nn.Conv2d(1, odim, 3, 2)`` of a ``Conv2dSubsampling4`` (its ``weight`` / ``bias`` are used as
    they are, so checkpoints load through the original module); ``global_cmvn``: a ``GlobalCMVN`` (reference or
    ``openeat_b200.cmvn``) or None.  Inference only: the kernel has no backward."""

    def __init__(self, conv, global_cmvn=None):
        super().__init__()
        assert conv.in_channels == 1 and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (2, 2)
        assert tuple(conv.padding) == (0, 0) and tuple(conv.dilation) == (1, 1)
        self.conv = conv
        self.global_cmvn = global_cmvn

    @torch.no_grad()
    def forward(self, xs):
        if not xs.is_cuda or xs.dtype != torch.float32:
            raise FrontendError('CmvnConvSubsample runs on fp32 CUDA tensors only (no CPU fallback)')
        lib = _lib.load()
        xs = xs.contiguous()
        B, T, F = xs.shape
        odim = self.conv.out_channels
        w = self.conv.weight.detach().to(xs.device, torch.float32).contiguous()
        b = self.conv.bias.detach().to(xs.device, torch.float32).contiguous() if self.conv.bias is not None else None
        mean = istd = None
        if self.global_cmvn is not None:
            mean = self.global_cmvn.mean.to(xs.device, torch.float32).contiguous()
            if getattr(self.global_cmvn, 'norm_var', True):
                istd = self.global_cmvn.istd.to(xs.device, torch.float32).contiguous()
        T1, F1 = (T - 3) // 2 + 1 if T >= 3 else 0, (F - 3) // 2 + 1
        y = torch.empty((B, odim, T1, F1), dtype=torch.float32, device=xs.device)

        def p(t):
            return ctypes.c_void_p(t.data_ptr()) if t is not None else None
        check(lib.oe_cmvn_conv_subsample(p(xs), F, B, T, F, p(mean), p(istd), p(w), p(b), odim, p(y),
                                         ctypes.c_void_p(torch.cuda.current_stream(xs.device).cuda_stream)))
        return y
