"""XNOR-Net ops -- surface of QuantTorch/functions/xnor_connect.py.

Reference quirks kept on purpose:
  * the activation op is sign(x) * mean(x, dim) -- the plain mean, torch.sign (0 -> 0)  (xnor_connect.py:20-28);
  * XNORDense scales by alpha = mean(|W|, DIM=0): one alpha per INPUT feature k, inside the contraction
    (xnor_connect.py:13,111-112); its `dim` argument is ignored.
The faithful product  y[b,o] = mu[b] * sum_k alpha[k] s_a[b,k] s_w[o,k]  is not a popcount; it runs on the
bf16 tensor-core route with the weights kept as 2 bit planes + alpha[k] in HBM.
"""
import torch

from .. import _engine as eng
from .. import _lib as L
from .. import _ops as ops
from .common import TaggingFunction, front

DIM = 0


def _quantOpXnor(dim=1):
    class _QuantXNOR(TaggingFunction):
        @staticmethod
        def forward(ctx, input):
            if dim == 1 and input.dim() == 2:
                # fused device pass: row mean (fp64 accumulate) + sign + fp16 (or bf16) codes for the next layer
                full = eng.want_fp32_result(input)
                y, tag = ops.quant_act(input, L.Q_XNOR_ROW, want_y=full, codes_kind=eng.xnor_codes_kind(),
                                       want_row_scale=True, kind="xnor")
                ctx.save_for_backward(input, tag.row_scale)
                TaggingFunction._leave(tag)
                return y if full else eng.placeholder_like(input)
            # dim 0 / -1 reduce over the batch: a column/global reduction followed by one elementwise product
            # (not used by the sharded configs -- it would need an all-reduce)
            ops.require_cuda(input, "input")
            mean = torch.mean(input) if dim < 0 else torch.mean(input, dim)
            ctx.save_for_backward(input, mean)
            if dim < 0:
                return torch.sign(input) * mean
            form_mean = {0: (1, -1), 1: (-1, 1)}[dim]
            return torch.sign(input) * mean.view(form_mean)

        @staticmethod
        def backward(ctx, grad_outputs):
            input, mean = ctx.saved_tensors
            sgn_input = torch.sign(input)
            if dim < 0:
                return sgn_input * torch.mean(grad_outputs * sgn_input) + grad_outputs * mean
            form_mean = {0: (1, -1), 1: (-1, 1)}[dim]
            return (sgn_input * torch.mean(grad_outputs * sgn_input, dim, keepdim=True)
                    + grad_outputs * mean.view(form_mean).expand(input.size()))
    return _QuantXNOR


_op_cache = {}


def _op(dim):
    if dim not in (-1, 0, 1):
        raise RuntimeError(" Please use a correct dim between -1, 0, 1")
    if dim not in _op_cache:
        _op_cache[dim] = _quantOpXnor(dim)
    return _op_cache[dim]


def nnQuantXnor(dim=1):
    """Module form of QuantXnor (xnor_connect.py:40-52)."""
    m = front(_op(dim))
    m._qt_spec = ("xnor", dim) if dim == 1 else None
    return m


def QuantXnor(input, dim=1):
    """sign(input) * mean(input, dim) for 2-D inputs (xnor_connect.py:54-66)."""
    return _op(dim).apply(input)


def _quantOpXnor2d(kernel_size, stride=1, padding=1, dilation=1, groups=1, form="NCHW"):
    """As in the reference (xnor_connect.py:69-89) this validates its arguments and returns None: the 2-D
    activation op was never finished upstream (its backward raises NotImplementedError)."""
    if form not in ["NHWC", "NCHW"]:
        raise RuntimeError("Input form insupported ")
    if type(kernel_size) != int:
        raise RuntimeError("Only int kernel_size supported (square kernel)")
    return None


def xnor_pack(weight):
    """2 bit planes (nz, sign) + alpha[k] = mean(|W|, 0) for a [out, in] weight."""
    return ops.pack_weight(weight.detach(), "xnor")


def xnor_conv_pack(weight, dim=(0, 1)):
    """XNORConv2d weights: alpha = mean(|W|, dim=[0,1], keepdim) is one value per filter tap (xnor_connect.py:140)."""
    O, Cg, kh, kw = weight.shape
    w = weight.detach()
    if sorted(d % 4 for d in dim) == [0, 1]:
        a_tap = ops.col_absmean(w.reshape(O * Cg, kh * kw))           # [kh*kw]
        alpha_k = a_tap.repeat_interleave(Cg).contiguous()            # column (kh, kw, c) -> alpha[kh, kw]
        return ops.pack_weight(ops.conv_weight_2d(w), "xnor", alpha=alpha_k)
    mean_weight = torch.mean(torch.abs(w), list(dim), keepdim=True)   # general `dim`: real-valued weight operand
    return ops.pack_real_weight(ops.conv_weight_2d(torch.sign(w) * mean_weight))


def XNORDense(dim=[0, 1]):
    """Dense op, weights binarised as sign(W) * mean(|W|, DIM=0)  (xnor_connect.py:93-132)."""
    class _XNORDense(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias=None, *cached_pack):
            # `cached_pack`: optional pre-built WeightPack (the LinearXNOR layer passes its eval-mode cache)
            ctx.n_extra = len(cached_pack)
            pack = cached_pack[0] if cached_pack and cached_pack[0] is not None else xnor_pack(weight)
            ctx.save_for_backward(input, weight, pack.alpha.view(1, -1), bias)
            return eng.linear(input, pack, bias)

        @staticmethod
        def backward(ctx, grad_output):
            input, weight, mean, bias = ctx.saved_tensors
            weight_q = torch.sign(weight) * mean
            gi = gw = gb = None
            if ctx.needs_input_grad[0]:
                gi = eng.grad_input_linear(grad_output, weight_q)
            if ctx.needs_input_grad[1]:
                t = eng.grad_weight_linear(grad_output, input)
                gw = mean * t + torch.sign(weight) * torch.mean(t * torch.sign(weight), DIM, keepdim=True)
            if bias is not None and ctx.needs_input_grad[2]:
                gb = grad_output.sum(0).squeeze(0)
            return (gi, gw, gb) + (None,) * ctx.n_extra
    return _XNORDense


def XNORConv2d(dim=[0, 1], quant_input=False, stride=1, padding=1, dilation=1, groups=1):
    """Conv op, weights sign(W) * mean(|W|, dim)  (xnor_connect.py:135-169)."""
    class _XNORConv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias=None, *cached_pack):
            ctx.n_extra = len(cached_pack)
            mean_weight = torch.mean(torch.abs(weight), dim, keepdim=True)
            if quant_input:
                input = torch.sign(input) * torch.mean(torch.abs(input), 1, keepdim=True)
            ctx.save_for_backward(input, weight, mean_weight, bias)
            pack = cached_pack[0] if cached_pack and cached_pack[0] is not None else xnor_conv_pack(weight, dim)
            return eng.conv2d(input, pack, bias, tuple(weight.shape), stride, padding, dilation, groups)

        @staticmethod
        def backward(ctx, grad_output):
            input, weight, mean, bias = ctx.saved_tensors
            weight_b = torch.sign(weight) * mean
            gi = gw = gb = None
            if ctx.needs_input_grad[0]:
                gi = eng.grad_input_conv2d(input.size(), weight_b, grad_output, stride=stride, padding=padding,
                                                dilation=dilation, groups=groups)
            if ctx.needs_input_grad[1]:
                t = eng.grad_weight_conv2d(input, weight.shape, grad_output, stride=stride, padding=padding,
                                                dilation=dilation, groups=groups)
                gw = mean * t + torch.sign(weight) * torch.mean(t * torch.sign(weight), DIM, keepdim=True)
            if bias is not None and ctx.needs_input_grad[2]:
                gb = grad_output.sum((0, 2, 3))
            return (gi, gw, gb) + (None,) * ctx.n_extra
    return _XNORConv2d
