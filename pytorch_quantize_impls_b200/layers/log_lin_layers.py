"""LogLin layers -- surface of QuantTorch/layers/log_lin_layers.py.

Quantized weights are sign * 2^e (log) or multiples of 2^(fsr-bit_width) (lin): both are exact in bf16, so the
contraction runs on the bf16 tensor-core route with the quantized values in the hi plane."""
import torch

from .. import _ops as ops
from ..functions import log_lin_connect
from .common import QuantLayerMixin, check_convert


class _LogLinMixin(QuantLayerMixin):
    def _weight_op(self, w):
        return self.weight_op.forward(w)

    def _make_pack(self, w):
        with torch.no_grad():
            wq = self.weight_op.forward(w.detach())
        if 1 <= self.bit_width <= 6:      # int8 codes in HBM (8 bits per weight; exact int8 / bf16 operands)
            return ops.pack_loglin_weight(self._w2d(wq), self._dtype, self.fsr, self.bit_width)
        return ops.pack_real_weight(self._w2d(wq))

    def _weight_op_host(self, w):
        if self._dtype == "lin":                                             # log_lin_connect.py:61-68
            if self.bit_width == 32:
                return w.clone()
            step = 2.0 ** (self.fsr - self.bit_width)
            return torch.sign(w) * torch.clamp(torch.round(torch.abs(w) / step) * step, 0, 2 ** self.fsr)
        e = torch.clamp(torch.round(torch.log2(torch.abs(w))), self.fsr - 2 ** self.bit_width, self.fsr)
        return torch.sign(w) * torch.pow(torch.ones_like(w) * 2, e)         # log_lin_connect.py:29-32

    def clamp(self):
        self.weight.data.clamp_(-1 * 2 ** (self.fsr), 2 ** (self.fsr))

    def _reset_loglin(self):
        torch.nn.init.uniform_(self.weight, 2 ** (self.fsr - self.bit_width), 2 ** (self.fsr))
        self.weight.data.mul_((torch.rand_like(self.weight) < 0.5).type(self.weight.dtype) * 2 - 1)
        if self.bias is not None:
            self.bias.data.zero_()


class LinearQuant(_LogLinMixin, torch.nn.Linear):
    """y = x . Q(W)^T + b; as in the reference the weight op is applied on EVERY forward, also in eval mode
    (log_lin_layers.py:40-42), where it re-quantizes the already-quantized weights (idempotent)."""

    @staticmethod
    def convert(other, dtype="lin", fsr=7, bit_width=3):
        check_convert(other, torch.nn.Linear, "torch.nn.Linear")
        return LinearQuant(other.in_features, other.out_features, False if other.bias is None else True, dtype=dtype,
                           fsr=fsr, bit_width=bit_width)

    def __init__(self, in_features, out_features, bias=True, dtype="lin", fsr=7, bit_width=3):
        self.bit_width = bit_width
        self.fsr = fsr
        self._dtype = dtype
        torch.nn.Linear.__init__(self, in_features, out_features, bias=bias)
        self.weight_op = log_lin_connect.nnQuant(dtype=dtype, fsr=fsr, bit_width=bit_width, with_sign=True,
                                                 lin_back=True, _emit_codes=False)

    def reset_parameters(self):
        self._reset_loglin()


class QuantConv2d(_LogLinMixin, torch.nn.Conv2d):
    """conv2d(x, Q(W)) + b (log_lin_layers.py:45-93)."""
    _is_conv = True

    @staticmethod
    def convert(other, fsr=7, bit_width=3, dtype="lin"):
        check_convert(other, torch.nn.Conv2d, "torch.nn.Conv2d")
        return QuantConv2d(other.in_channels, other.out_channels, other.kernel_size, stride=other.stride,
                           padding=other.padding, dilation=other.dilation, groups=other.groups,
                           bias=False if other.bias is None else True, fsr=fsr, bit_width=bit_width, dtype=dtype)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 fsr=7, bit_width=3, dtype="lin"):
        self.fsr = fsr
        self.bit_width = bit_width
        self._dtype = dtype
        torch.nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                 dilation=dilation, groups=groups, bias=bias)
        self.weight_op = log_lin_connect.nnQuant(dtype=dtype, fsr=fsr, bit_width=bit_width, with_sign=True,
                                                 lin_back=True, _emit_codes=False)

    def reset_parameters(self):
        if self.bit_width == 32:
            super().reset_parameters()
        self._reset_loglin()
