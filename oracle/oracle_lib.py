"""CPU twin namespace for pytorch_quantize_impls_b200.nets builders: the same class names as the product layers,
implemented with the oracle (fake-quant + fp32 F.linear / F.conv2d) -- i.e. the reference's path.  TEST INFRASTRUCTURE: imported by tests/ and by
bench.py's CPU-baseline / --impl reference legs only."""
import torch
from torch import nn

import quanttorch_oracle as O


class _Fn(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


def BinaryConnect(stochastic=False):
    return _Fn(O.binary_det)


def TernaryConnect(stochastic=False):
    return _Fn(O.ternary_det)


def nnDorefaQuant(bit_width=3):
    return _Fn(lambda x: O.dorefa_quantize(x, bit_width))


def nnQuantXnor(dim=1):
    return _Fn(lambda x: O.xnor_act(x, dim))


def _lin(fn):
    class L(nn.Linear):
        def __init__(self, i, o, bias=True, **kw):
            super().__init__(i, o, bias=bias)
            self.kw = kw

        def forward(self, x):
            return fn(x, self.weight, self.bias, **self.kw)
    return L


def _conv(fn):
    class C(nn.Conv2d):
        def __init__(self, i, o, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True, **kw):
            super().__init__(i, o, kernel_size, stride=stride, padding=padding, dilation=dilation, groups=groups, bias=bias)
            self.kw = kw

        def forward(self, x):
            return fn(x, self.weight, self.bias, stride=self.stride, padding=self.padding, dilation=self.dilation,
                      groups=self.groups, **self.kw)
    return C


LinearBin, BinConv2d = _lin(O.linear_bin), _conv(O.conv_bin)
LinearTer, TerConv2d = _lin(O.linear_ter), _conv(O.conv_ter)
LinearDorefa, DorefaConv2d = _lin(O.linear_dorefa), _conv(O.conv_dorefa)
LinearXNOR, XNORConv2d = _lin(O.linear_xnor), _conv(O.conv_xnor)
LinearQuant, QuantConv2d = _lin(O.linear_loglin), _conv(O.conv_loglin)
