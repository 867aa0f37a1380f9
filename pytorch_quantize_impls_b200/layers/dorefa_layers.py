"""DoReFa layers -- surface of QuantTorch/layers/dorefa_layers.py."""
import torch

from .. import _ops as ops
from ..functions import dorefa_connect
from .common import QuantLayerMixin, check_convert


class _DorefaMixin(QuantLayerMixin):
    def _init_dorefa(self, bit_width):
        self.bit_width = bit_width
        self.weight_op = dorefa_connect.nnQuantWeight(bit_width=bit_width)

    def _weight_op(self, w):
        return self.weight_op.forward(w)

    def _make_pack(self, w):
        w2 = self._w2d(w.detach())
        if 1 <= self.bit_width <= 8:
            return ops.pack_weight(w2, "dorefa", self.bit_width)     # k-bit codes (+ E / 1/n column scale)
        with torch.no_grad():                                        # k == 32 (or 9..16): real-valued operand
            return ops.pack_real_weight(self._w2d(self.weight_op.forward(w.detach())))


    def _weight_op_host(self, w):
        k = self.bit_width                                                   # dorefa_connect.py:99-111
        if k == 32:
            return w.clone()
        one = torch.ones_like(w)
        if k == 1:
            return torch.where(w < 0, -one, one) * torch.mean(torch.abs(w))
        if torch.max(torch.abs(w)) == 0.0:
            return torch.zeros_like(w)
        t = torch.tanh(w)
        t = t / (2 * torch.max(torch.abs(t))) + 0.5
        n = torch.pow(one * 2, k) - 1
        return 2 * ((1 / n) * torch.round(n * t)) - 1


class LinearDorefa(_DorefaMixin, torch.nn.Linear):
    """y = x . quantize_w(W)^T + b with k-bit weights (dorefa_layers.py:11-45)."""

    @staticmethod
    def convert(other, bit_width=3):
        check_convert(other, torch.nn.Linear, "torch.nn.Linear")
        return LinearDorefa(other.in_features, other.out_features, False if other.bias is None else True,
                            bit_width=bit_width)

    def __init__(self, in_features, out_features, bias=True, bit_width=3):
        torch.nn.Linear.__init__(self, in_features, out_features, bias=bias)
        self._init_dorefa(bit_width)

    def extra_repr(self):
        return "bit_width = {}".format(self.bit_width)


class DorefaConv2d(_DorefaMixin, torch.nn.Conv2d):
    """conv2d(x, quantize_w(W)) + b with k-bit weights (dorefa_layers.py:48-82)."""
    _is_conv = True

    @staticmethod
    def convert(other, bit_width=3):
        check_convert(other, torch.nn.Conv2d, "torch.nn.Conv2d")
        return DorefaConv2d(other.in_channels, other.out_channels, other.kernel_size, stride=other.stride,
                            padding=other.padding, dilation=other.dilation, groups=other.groups,
                            bias=False if other.bias is None else True, bit_width=bit_width)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 bit_width=3):
        torch.nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                 dilation=dilation, groups=groups, bias=bias)
        self._init_dorefa(bit_width)
