"""BinaryConnect / BinaryNet ops -- surface of QuantTorch/functions/binary_connect.py."""
import warnings as _warnings

import torch

from .. import _engine as eng
from .. import _lib as L
from .. import _ops as ops
from .common import TaggingFunction, front, safeSign, ste_clip


class BinaryConnectDeterministic(TaggingFunction):
    """r_b = sign(r) (0 -> +1); backward 1_{|r|<=1.001}.  binary_connect.py:14-38.

    Besides the fp32 +-1 tensor the forward emits int8 codes and packed sign bits in the same pass; the next
    LinearBin / BinConv2d / LinearTer / LinearDorefa contracts on those."""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        full = eng.want_fp32_result(input)
        y, tag = ops.quant_act(input, L.Q_SIGN, want_y=full, codes_kind=eng.int_codes_kind(input),
                               want_bits=eng.want_sign_bits(input), kind="sign")
        TaggingFunction._leave(tag)
        return y if full else eng.placeholder_like(input)

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        return ste_clip(grad_output, input)


class BinaryConnectStochastic(torch.autograd.Function):
    """+1 with probability hardsigmoid(r); binary_connect.py:42-71.  Device Philox stream (torch.rand_like),
    so it is statistically, not bitwise, comparable with the CPU reference (as the reference's own test does)."""

    @staticmethod
    def forward(ctx, input):
        ops.require_cuda(input, "input")
        ctx.save_for_backward(input)
        z = torch.rand_like(input, requires_grad=False)
        p = (torch.clamp(input, -1, 1) + 1) / 2
        return -1.0 + 2.0 * (z < p).float()

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        return ste_clip(grad_output, input)


def BinaryConnect(stochastic=False):
    """nn.Module wrapping the binarization op (binary_connect.py:74-83)."""
    m = front(BinaryConnectStochastic if stochastic else BinaryConnectDeterministic)
    m._qt_spec = None if stochastic else ("sign", 1)
    return m


def _sign_pack(weight):
    return ops.pack_weight(ops.conv_weight_2d(weight.detach()), "sign")


class BinaryDense(TaggingFunction):
    """y = x . sign(W)^T + b with explicit backward, binary_connect.py:86-112."""

    @staticmethod
    def forward(ctx, input, weight, bias=None):
        ctx.save_for_backward(input, weight, bias)
        return eng.linear(input, _sign_pack(weight), bias)

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, bias = ctx.saved_tensors
        weight_b = safeSign(weight)
        grad_input = grad_weight = grad_bias = None
        if ctx.needs_input_grad[0]:
            grad_input = grad_output.mm(weight_b)
        if ctx.needs_input_grad[1]:
            grad_weight = grad_output.t().mm(input)
        if bias is not None and ctx.needs_input_grad[2]:
            grad_bias = grad_output.sum(0).squeeze(0)
        return grad_input, grad_weight, grad_bias


def BinaryConv2d(stride=1, padding=1, dilation=1, groups=1):
    """DEPRECATED functional conv (binary_connect.py:116-153); kept for surface parity."""
    _warnings.warn("Deprecated conv op !", DeprecationWarning, stacklevel=2)

    class _BinaryConv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias=None):
            ctx.save_for_backward(input, weight, bias)
            return eng.conv2d(input, _sign_pack(weight), bias, tuple(weight.shape), stride, padding, dilation, groups)

        @staticmethod
        def backward(ctx, grad_output):
            input, weight, bias = ctx.saved_tensors
            weight_b = safeSign(weight)
            gi = gw = gb = None
            if ctx.needs_input_grad[0]:
                gi = eng.grad_input_conv2d(input.size(), weight_b, grad_output, stride=stride, padding=padding,
                                                dilation=dilation, groups=groups)
            if ctx.needs_input_grad[1]:
                gw = eng.grad_weight_conv2d(input, weight.shape, grad_output, stride=stride, padding=padding,
                                                 dilation=dilation, groups=groups)
            if bias is not None and ctx.needs_input_grad[2]:
                gb = grad_output.sum((0, 2, 3))
            return (gi, gw, gb) if bias is not None else (gi, gw)

    return _BinaryConv2d
