"""Ternary connect ops -- surface of QuantTorch/functions/terner_connect.py."""
import warnings

import torch

from .. import _engine as eng
from .. import _lib as L
from .. import _ops as ops
from .common import TaggingFunction, front, safeSign, ste_clip


class TernaryConnectDeterministic(TaggingFunction):
    """x >= .5 -> 1, -.5 <= x < .5 -> 0, x < -.5 -> -1 (ties as the reference: +.5 -> 1, -.5 -> 0);
    backward 1_{|x|<=1.001}.  terner_connect.py:12-34."""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        full = eng.want_fp32_result(input)
        y, tag = ops.quant_act(input, L.Q_TERNARY, want_y=full, codes_kind=eng.int_codes_kind(input), kind="ternary")
        TaggingFunction._leave(tag)
        return y if full else eng.placeholder_like(input)

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        return ste_clip(grad_output, input)


class TernaryConnectStochastic(torch.autograd.Function):
    """s - s * 1[z > |x|], z ~ U[0,1).  terner_connect.py:37-63 (device RNG; statistical parity only)."""

    @staticmethod
    def forward(ctx, input):
        ops.require_cuda(input, "input")
        ctx.save_for_backward(input)
        sign = safeSign(input)
        z = torch.rand_like(input, requires_grad=False)
        return sign - sign * (z > torch.abs(input)).to(input.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        return ste_clip(grad_output, input)


def TernaryConnect(stochastic=False):
    """nn.Module wrapping the ternary op (terner_connect.py:67-75)."""
    m = front(TernaryConnectStochastic if stochastic else TernaryConnectDeterministic)
    m._qt_spec = None if stochastic else ("ternary", 2)
    return m


def _functional_ternary(weight, stochastic):
    # the *functional* ops use torch.sign, not safeSign (terner_connect.py:83-90): +-0.5 -> +-0.5, 0 -> 0.
    # They are off the layer path; the weight transform is tiny next to the contraction and runs as torch ops.
    sign = torch.sign(weight)
    if stochastic:
        z = torch.rand_like(weight, requires_grad=False)
        return sign - torch.sign(z - torch.abs(weight))
    return (sign + torch.sign(weight - 0.5 * sign)) / 2


def TernaryDense(stochastic=False):
    """Linear op with ternary weights and explicit backward, terner_connect.py:78-108."""
    class _TernaryDense(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias=None):
            weight_t = _functional_ternary(weight, stochastic)
            ctx.save_for_backward(input, weight, weight_t, bias)
            return eng.linear(input, ops.pack_real_weight(weight_t), bias)

        @staticmethod
        def backward(ctx, grad_output):
            input, weight, weight_t, bias = ctx.saved_tensors
            gi = gw = gb = None
            if ctx.needs_input_grad[0]:
                gi = grad_output.mm(weight_t)
            if ctx.needs_input_grad[1]:
                gw = grad_output.t().mm(input)
            if bias is not None and ctx.needs_input_grad[2]:
                gb = grad_output.sum(0).squeeze(0)
            return gi, gw, gb
    return _TernaryDense


def TernaryConv2d(stochastic=True, stride=1, padding=1, dilation=1, groups=1):
    """DEPRECATED functional conv, terner_connect.py:113-153."""
    warnings.warn("Deprecated conv op !", DeprecationWarning, stacklevel=2)

    class _TernaryConv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias=None):
            weight_t = _functional_ternary(weight, stochastic)
            ctx.save_for_backward(input, weight, weight_t, bias)
            pack = ops.pack_real_weight(ops.conv_weight_2d(weight_t))
            return eng.conv2d(input, pack, bias, tuple(weight.shape), stride, padding, dilation, groups)

        @staticmethod
        def backward(ctx, grad_output):
            input, weight, weight_t, bias = ctx.saved_tensors
            gi = gw = gb = None
            if ctx.needs_input_grad[0]:
                gi = eng.grad_input_conv2d(input.size(), weight_t, grad_output, stride=stride, padding=padding,
                                                dilation=dilation, groups=groups)
            if ctx.needs_input_grad[1]:
                gw = eng.grad_weight_conv2d(input, weight.shape, grad_output, stride=stride, padding=padding,
                                                 dilation=dilation, groups=groups)
            if bias is not None and ctx.needs_input_grad[2]:
                gb = grad_output.sum((0, 2, 3))
            return (gi, gw, gb) if bias is not None else (gi, gw)
    return _TernaryConv2d
