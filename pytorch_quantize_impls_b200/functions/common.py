"""safeSign / front -- mirrors QuantTorch/functions/common.py:4-31 on top of the CUDA quantizer kernels."""
import threading

import numpy as np
import torch

from .. import _engine as eng
from .. import _lib as L
from .. import _ops as ops

_tls = threading.local()


def _f32(v):
    """Python float holding exactly the fp32 value v (e.g. fl(1/n))."""
    return float(np.float32(v))


def safeSign(tensor):
    """sign(x) with 0 -> +1 (also -0.0 and NaN), QuantTorch/functions/common.py:4-7.  New tensor, input untouched."""
    y, _ = ops.quant_act(tensor, L.Q_SIGN, want_y=True)
    return y


class TaggingFunction(torch.autograd.Function):
    """autograd.Function whose forward may leave a low-bit operand (ActCodes) for the consumer layer.

    `forward` stores the operand in a thread-local slot; `apply` attaches it to the returned tensor as
    `_qt_codes` so that LinearBin / BinConv2d / ... can contract on the codes instead of re-reading fp32."""

    @classmethod
    def apply(cls, *args, **kwargs):
        _tls.pending = None
        eng._apply_grad_mode.value = torch.is_grad_enabled()     # grad mode is forced off inside forward()
        try:
            out = super().apply(*args, **kwargs)
        finally:
            eng._apply_grad_mode.value = None
        tag = getattr(_tls, "pending", None)
        _tls.pending = None
        if tag is not None and isinstance(out, torch.Tensor):
            eng.attach_tag(out, tag)
        return out

    @staticmethod
    def _leave(tag):
        _tls.pending = tag


def front(claaz):
    """Module proxy of an autograd.Function class, QuantTorch/functions/common.py:10-20."""
    class fronteur(torch.nn.Module):
        def forward(self, x):
            return claaz.apply(x)
    return fronteur()


def front2(claaz):
    """QuantTorch/functions/common.py:23-31."""
    class fronteur(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.core = claaz

        def forward(self, x):
            return self.core.apply(x)
    return fronteur()


def ste_clip(grad_output, x):
    """d/dx = 1_{|x| <= 1.001}: binary_connect.py:30-38, terner_connect.py:29-34."""
    if grad_output.is_cuda and x.is_cuda and grad_output.dtype == torch.float32 and x.dtype == torch.float32:
        return ops.ste_clip(grad_output, x, 1.001)       # one pass (qt_ste_clip)
    g = grad_output.clone()
    g[torch.abs(x) > 1.001] = 0
    return g
