"""BinaryNet layers -- surface of QuantTorch/layers/binary_layers.py."""
from math import sqrt as _sqrt

import torch

from .. import _ops as ops
from ..functions import binary_connect
from .common import QuantLayerMixin, check_convert


class _BinMixin(QuantLayerMixin):
    def _init_bin(self, deterministic):
        self.deterministic = deterministic
        self.bin_op = (binary_connect.BinaryConnectDeterministic if deterministic
                       else binary_connect.BinaryConnectStochastic)

    def _weight_op(self, w):
        return self.bin_op.apply(w)

    def _make_pack(self, w):
        if self.deterministic:
            return ops.pack_weight(self._w2d(w.detach()), "sign")
        # stochastic binarisation draws new +-1 weights each call; pack the drawn signs
        return self._make_pack_of_sample(self.bin_op.apply(w.detach()))

    def _make_pack_of_sample(self, wq):
        return ops.pack_weight(self._w2d(wq.detach()), "sign")      # sign(+-1) = +-1

    def _weight_op_host(self, w):
        one = torch.ones_like(w)
        if self.deterministic:
            return torch.where(w < 0, -one, one)                             # 0, -0.0, NaN -> +1 (functions/common.py:4-7)
        p = (torch.clamp(w, -1, 1) + 1) / 2                                  # binary_connect.py:57-61
        return torch.where(torch.rand_like(w) < p, one, -one)


class LinearBin(_BinMixin, torch.nn.Linear):
    """y = x . sign(W)^T + b   (binary_layers.py:7-46).  With a BinaryConnect() activation upstream the product is
    a 1-bit x 1-bit contraction with exact integer accumulators."""

    @staticmethod
    def convert(other, deterministic=True):
        check_convert(other, torch.nn.Linear, "torch.nn.Linear")
        return LinearBin(other.in_features, other.out_features, False if other.bias is None else True, deterministic)

    def __init__(self, in_features, out_features, bias=True, deterministic=True):
        torch.nn.Linear.__init__(self, in_features, out_features, bias=bias)
        self._init_bin(deterministic)

    def reset_parameters(self):
        self.weight.data.normal_(0, 1 * (_sqrt(1. / self.in_features)))
        if self.bias is not None:
            self.bias.data.zero_()

    def clamp(self):
        self.weight.data.clamp_(-1, 1)
        if self.bias is not None:
            self.bias.data.clamp_(-1, 1)


class BinConv2d(_BinMixin, torch.nn.Conv2d):
    """conv2d(x, sign(W)) + b   (binary_layers.py:48-106)."""
    _is_conv = True

    @staticmethod
    def convert(other, deterministic=True):
        check_convert(other, torch.nn.Conv2d, "torch.nn.Conv2d")
        return BinConv2d(other.in_channels, other.out_channels, other.kernel_size, stride=other.stride,
                         padding=other.padding, dilation=other.dilation, groups=other.groups,
                         bias=False if other.bias is None else True, deterministic=deterministic)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 deterministic=True):
        torch.nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                 dilation=dilation, groups=groups, bias=bias)
        self._init_bin(deterministic)

    def clamp(self):
        """Clamp real weights to [-1, 1] (weights only, binary_layers.py:81-85)."""
        self.weight.data.clamp_(-1, 1)
