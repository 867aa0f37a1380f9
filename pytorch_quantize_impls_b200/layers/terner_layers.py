"""Ternary layers -- surface of QuantTorch/layers/terner_layers.py."""
from math import sqrt

import torch

from .. import _ops as ops
from ..functions import terner_connect
from .common import QuantLayerMixin, check_convert


class _TerMixin(QuantLayerMixin):
    def _init_ter(self, deterministic):
        self.deterministic = deterministic
        self.ter_op = (terner_connect.TernaryConnectDeterministic if deterministic
                       else terner_connect.TernaryConnectStochastic)

    def _weight_op(self, w):
        return self.ter_op.apply(w)

    def _make_pack(self, w):
        w2 = self._w2d(w.detach())
        if self.deterministic:
            return ops.pack_weight(w2, "ternary")
        # stochastic: the drawn values are already in {-1, 0, 1}; the deterministic packer maps them to themselves
        return self._make_pack_of_sample(self.ter_op.apply(w.detach()))

    def _make_pack_of_sample(self, wq):
        return ops.pack_weight(self._w2d(wq.detach()), "ternary")

    def _weight_op_host(self, w):
        one = torch.ones_like(w)
        s = torch.where(w < 0, -one, one)
        if self.deterministic:
            return (s + torch.where(w - 0.5 * s < 0, -one, one)) / 2        # terner_connect.py:24-27
        return s - s * (torch.rand_like(w) > torch.abs(w)).to(w.dtype)       # terner_connect.py:50-55


class LinearTer(_TerMixin, torch.nn.Linear):
    """y = x . ter(W)^T + b, W in {-1, 0, 1} stored as two bit planes (terner_layers.py:10-51)."""

    @staticmethod
    def convert(other, dtype="lin", deterministic=True):
        check_convert(other, torch.nn.Linear, "torch.nn.Linear")
        return LinearTer(other.in_features, other.out_features, False if other.bias is None else True,
                         deterministic=deterministic)

    def __init__(self, in_features, out_features, bias=True, deterministic=True):
        torch.nn.Linear.__init__(self, in_features, out_features, bias=bias)
        self._init_ter(deterministic)

    def reset_parameters(self):
        self.weight.data.normal_(0, 1 * (sqrt(1. / self.in_features)))
        if self.bias is not None:
            self.bias.data.zero_()

    def clamp(self):
        self.weight.data.clamp_(-1, 1)
        if self.bias is not None:
            self.bias.data.clamp_(-1, 1)


class TerConv2d(_TerMixin, torch.nn.Conv2d):
    """conv2d(x, ter(W)) + b (terner_layers.py:54-92)."""
    _is_conv = True

    @staticmethod
    def convert(other, deterministic=True):
        check_convert(other, torch.nn.Conv2d, "torch.nn.Conv2d")
        return TerConv2d(other.in_channels, other.out_channels, other.kernel_size, stride=other.stride,
                         padding=other.padding, dilation=other.dilation, groups=other.groups,
                         bias=False if other.bias is None else True, deterministic=deterministic)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 deterministic=True):
        torch.nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                 dilation=dilation, groups=groups, bias=bias)
        self._init_ter(deterministic)

    def clamp(self):
        self.weight.data.clamp_(-1, 1)
