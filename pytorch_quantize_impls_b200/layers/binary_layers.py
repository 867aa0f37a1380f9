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
            return ops.pack_weight(ops.conv_weight_2d(w.detach()), "sign")
        # stochastic binarisation draws new +-1 weights each call; pack the drawn signs
        return ops.pack_weight(ops.conv_weight_2d(self.bin_op.apply(w.detach())), "sign")


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


class ShiftNormBatch1d(torch.nn.Module):
    """Shift-based batch norm, binary_layers.py:110-129 (off the measured path; torch ops)."""
    __constants__ = ['momentum', 'eps']

    def __init__(self, in_dim, eps=1e-5, momentum=0.1):
        super().__init__()
        self.in_features = in_dim
        self.weight = torch.nn.Parameter(torch.Tensor(self.in_features))
        self.bias = torch.nn.Parameter(torch.Tensor(self.in_features))
        self.register_buffer('running_mean', torch.zeros(self.in_features))
        self.register_buffer('running_var', torch.ones(self.in_features))
        self.eps = eps
        self.momentum = momentum

    def forward(self, x):
        self.running_mean = (1 - self.momentum) * self.running_mean + self.momentum * torch.mean(x, 0).detach()
        d = x - self.running_mean
        self.running_var = (1 - self.momentum) * self.running_var + self.momentum * torch.mean(
            d * binary_connect.AP2(d), 0).detach()
        return binary_connect.ShiftBatch.apply(x, self.running_mean, self.running_var, self.weight, self.bias, self.eps)


class ShiftNormBatch2d(torch.nn.Module):
    """2-D shift-based batch norm, binary_layers.py:134-160 (off the measured path; torch ops)."""
    __constants__ = ['momentum', 'eps']

    def __init__(self, in_channels, eps=1e-5, momentum=0.1):
        super().__init__()
        self.in_features = in_channels
        self.weight = torch.nn.Parameter(torch.Tensor(self.in_features))
        self.bias = torch.nn.Parameter(torch.Tensor(self.in_features))
        self.register_buffer('running_mean', torch.zeros(self.in_features))
        self.register_buffer('running_var', torch.ones(self.in_features))
        self.eps = eps
        self.momentum = momentum

    @staticmethod
    def _tile(tensor, dim):
        return tensor.repeat(dim[0], dim[1], 1).transpose(2, 0)

    def forward(self, x):
        dim = x.size()[-2:]
        self.running_mean = (1 - self.momentum) * self.running_mean + self.momentum * torch.mean(x, [0, 2, 3]).detach()
        curr_mean = ShiftNormBatch2d._tile(self.running_mean, dim)
        d = x - curr_mean
        self.running_var = (1 - self.momentum) * self.running_var + self.momentum * torch.mean(
            d * binary_connect.AP2(d), [0, 2, 3]).detach()
        return binary_connect.ShiftBatch.apply(x, curr_mean, ShiftNormBatch2d._tile(self.running_var, dim),
                                               ShiftNormBatch2d._tile(self.weight, dim),
                                               ShiftNormBatch2d._tile(self.bias, dim), self.eps)
