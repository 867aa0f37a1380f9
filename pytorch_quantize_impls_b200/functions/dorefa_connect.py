"""DoReFa-Net ops -- surface of QuantTorch/functions/dorefa_connect.py."""
import warnings

import torch

from .. import _engine as eng
from .. import _lib as L
from .. import _ops as ops
from .common import TaggingFunction, _f32, front, safeSign


def _quantize_with_codes(x, bit_width, pre=None):
    """_quantize (dorefa_connect.py:11-25) on the device.  Returns (y, ActCodes or None).

    k == 1 -> safeSign (+ sign codes/bits); k == 32 -> x itself; else y = fl(1/n) * round(n x), no clamp,
    round-half-even, with integer codes c = round(n x) in an int8 lane (k <= 7) or uint8 lane (k == 8) and
    their row sums."""
    full = eng.want_fp32_result(x)
    if bit_width == 1:
        y, tag = ops.quant_act(x, L.Q_SIGN, want_y=full, codes_kind=eng.int_codes_kind(x), want_bits=eng.want_sign_bits(x), kind="sign",
                               pre=pre)
        return (y if full else eng.placeholder_like(x)), tag
    if bit_width == 32:
        ops.require_cuda(x, "input")
        return x, None
    if 2 <= bit_width <= 8:
        kind = L.CODES_I8 if bit_width <= 7 else L.CODES_U8
        if bit_width == 2:
            kind = eng.int_codes_kind(x, 2)      # codes 0..3 are exact e2m1 values
        y, tag = ops.quant_act(x, L.Q_DOREFA, bit_width=bit_width, want_y=full, codes_kind=kind,
                               want_row_sum=(x.dim() == 2), kind="dorefa", pre=pre)
        tag.scale = _f32(1.0) / _f32(2 ** bit_width - 1)
        tag.scale = _f32(tag.scale)
        return (y if full else eng.placeholder_like(x)), tag
    if 9 <= bit_width <= 16:                     # no 8-bit lane: fp32 result only
        return ops.quant_act(x, L.Q_DOREFA, bit_width=bit_width, want_y=True)
    raise RuntimeError("bit_width %r not supported (1..16 or 32)" % (bit_width,))


def _quantize(x, bit_width=3):
    """quantize_k(x) = round((2^k-1) x) / (2^k-1)   (dorefa_connect.py:11-25)."""
    return _quantize_with_codes(x, bit_width)[0]


def _make_quant_function(bit_width):
    class _Quant(TaggingFunction):
        @staticmethod
        def forward(ctx, input):
            y, tag = _quantize_with_codes(input, bit_width)
            TaggingFunction._leave(tag)
            if y is input:
                y = input.view_as(input)
            return y

        @staticmethod
        def backward(ctx, grad_ouput):
            return grad_ouput.clone()
    return _Quant


_quant_cache = {}


def _quant_fn(bit_width):
    if bit_width not in _quant_cache:
        _quant_cache[bit_width] = _make_quant_function(bit_width)
    return _quant_cache[bit_width]


def nnDorefaQuant(bit_width=3):
    """nn.Module with the k-bit activation quantizer inside; identity STE (dorefa_connect.py:28-45)."""
    m = front(_quant_fn(bit_width))
    m._qt_spec = ("dorefa", bit_width) if 1 <= bit_width <= 8 else None
    return m


def DorefaQuant(x, bit_width=3):
    """Functional k-bit activation quantizer (dorefa_connect.py:49-63)."""
    return _quant_fn(bit_width).apply(x)


class _ignore_factor_op(torch.autograd.Function):
    """input * const with the factor ignored by the gradient (dorefa_connect.py:66-79)."""

    @staticmethod
    def forward(ctx, input, const):
        return input * const

    @staticmethod
    def backward(ctx, grad_ouput):
        return (grad_ouput.clone() if ctx.needs_input_grad[0] else None), None


class _WeightQuantSTE(torch.autograd.Function):
    """Forward of nnQuantWeight for 2 <= k <= 8 as ONE fused device pass (two-pass global max|tanh W| +
    code generation); backward = the autograd chain of the reference expression
    2*quantize_k(tanh(W)/(2 max|tanh W|) + 1/2) - 1 with an identity STE through quantize_k."""

    @staticmethod
    def forward(ctx, w, bit_width):
        ctx.save_for_backward(w)
        p = ops.pack_weight(w.detach().reshape(w.shape[0], -1) if w.dim() > 1 else w.detach().reshape(1, -1),
                            "dorefa", bit_width, want_wq=True)
        wq = torch.where(p.stats[3] == 0, torch.zeros_like(p.wq), p.wq)   # all-zero guard, :106-107
        return wq.reshape(w.shape)

    @staticmethod
    def backward(ctx, g):
        w, = ctx.saved_tensors
        with torch.enable_grad():
            wd = w.detach().requires_grad_(True)
            t = torch.tanh(wd)
            out = 2 * (t / (2 * torch.max(torch.abs(t))) + 0.5) - 1
            gw, = torch.autograd.grad(out, wd, g)
        return gw, None


def nnQuantWeight(bit_width=3):
    """Module quantizing a layer's weights, dorefa_connect.py:82-113:
    k == 1: safeSign(W) * mean|W|;  k == 32: W;  else 2*quantize_k(tanh(W)/(2 max|tanh W|) + 1/2) - 1
    (zeros if W is all zeros)."""
    class _QuantWeight(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.bit_width = bit_width
            self.quant_op = nnDorefaQuant(bit_width)

        def forward(self, x):
            if self.bit_width == 1:
                E = torch.mean(torch.abs(x)).detach()
                return _ignore_factor_op.apply(self.quant_op(x), E)
            if self.bit_width == 32:
                return x
            if 2 <= self.bit_width <= 8:
                return _WeightQuantSTE.apply(x, self.bit_width)
            if torch.max(torch.abs(x)) == 0.0:
                return torch.zeros_like(x)
            weight = torch.tanh(x)
            weight = weight / (2 * torch.max(torch.abs(weight))) + 0.5
            return 2 * self.quant_op(weight) - 1
    return _QuantWeight()


def dorefa_pack(weight, bit_width):
    """k-bit HBM pack of a layer weight (rows = out features / channels)."""
    return ops.pack_weight(ops.conv_weight_2d(weight.detach()), "dorefa", bit_width)


def _functional_weight(weight, bit_width, max_abs):
    if bit_width == 1:
        return safeSign(weight) * torch.mean(torch.abs(weight)).detach()
    if bit_width == 32:
        return weight
    return 2 * _quantize(0.5 + torch.tanh(weight) / (2 * max_abs), bit_width=bit_width) - 1


def QuantDense(bit_width=3):
    """DEPRECATED fused dense op with explicit backward, dorefa_connect.py:116-155."""
    class _QuantDense(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias=None):
            max_abs = torch.max(torch.abs(torch.tanh(weight)))
            weight_q = _functional_weight(weight, bit_width, max_abs)
            output = eng.linear(input, ops.pack_real_weight(weight_q), bias)
            ctx.save_for_backward(input, weight, weight_q, max_abs, bias)
            return output

        @staticmethod
        def backward(ctx, grad_output):
            input, weight, weight_q, max_abs, bias = ctx.saved_tensors
            gi = gw = gb = None
            if ctx.needs_input_grad[0]:
                gi = grad_output.mm(weight_q)
            if ctx.needs_input_grad[1]:
                gw = grad_output.t().mm(input)
                if bit_width not in (1, 32):
                    gw = gw * (1 - torch.pow(torch.tanh(weight), 2)) / max_abs
            if bias is not None and ctx.needs_input_grad[2]:
                gb = grad_output.sum(0).squeeze(0)
            return gi, gw, gb
    return _QuantDense


def QuantConv2d(stride=1, padding=1, dilation=1, groups=1, bit_width=3):
    """DEPRECATED fused conv op, dorefa_connect.py:158-199 (normalises by tanh(max|W|), the same value)."""
    warnings.warn("Deprecated conv op !", DeprecationWarning, stacklevel=2)

    class _QuantConv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias=None):
            max_weight = torch.max(torch.abs(weight))
            weight_q = _functional_weight(weight, bit_width, torch.tanh(max_weight))
            ctx.save_for_backward(input, weight, weight_q, max_weight, bias)
            pack = ops.pack_real_weight(ops.conv_weight_2d(weight_q))
            return eng.conv2d(input, pack, bias, tuple(weight.shape), stride, padding, dilation, groups)

        @staticmethod
        def backward(ctx, grad_output):
            input, weight, weight_q, max_weight, bias = ctx.saved_tensors
            gi = gw = gb = None
            if ctx.needs_input_grad[0]:
                gi = eng.grad_input_conv2d(input.size(), weight_q, grad_output, stride=stride, padding=padding,
                                                dilation=dilation, groups=groups)
            if ctx.needs_input_grad[1]:
                gw = eng.grad_weight_conv2d(input, weight.shape, grad_output, stride=stride, padding=padding,
                                                 dilation=dilation, groups=groups)
                if 1 < bit_width < 32:
                    gw = gw * (1 - torch.pow(torch.tanh(weight), 2)) / torch.tanh(max_weight)
            if bias is not None and ctx.needs_input_grad[2]:
                gb = grad_output.sum((0, 2, 3))
            return (gi, gw, gb) if bias is not None else (gi, gw)
    return _QuantConv2d
