"""Miyashita lin/log fixed-point quantizers -- surface of QuantTorch/functions/log_lin_connect.py."""
import torch

from .. import _engine as eng
from .. import _lib as L
from .. import _ops as ops
from .common import TaggingFunction, front


def _device_quant(x, mode, fsr, bit_width, with_sign):
    return ops.quant_act(x, mode, bit_width=bit_width, fsr=fsr, with_sign=with_sign, want_y=True)[0]


def _device_quant_tagged(x, mode, fsr, bit_width, with_sign):
    """Quantizer pass that also leaves the low-bit operand for the next layer:
    lin -> int8 (signed) / uint8 codes value / step with scale = step  (exact integer product with Lin-coded weights);
    log -> the powers of two themselves in a bf16 lane (exact)."""
    if x.dim() not in (2, 4) or not x.is_cuda:
        return _device_quant(x, mode, fsr, bit_width, with_sign), None
    full = eng.want_fp32_result(x)
    if mode == L.Q_LIN and bit_width <= (6 if with_sign else 7):
        kind = L.CODES_I8 if with_sign else L.CODES_U8
        y, tag = ops.quant_act(x, mode, bit_width=bit_width, fsr=fsr, with_sign=with_sign, want_y=full, codes_kind=kind,
                               want_row_sum=(x.dim() == 2), kind="lin")
        tag.scale = float(2.0 ** (fsr - bit_width))
    elif mode == L.Q_LOG and x.dim() == 2:
        y, tag = ops.quant_act(x, mode, bit_width=bit_width, fsr=fsr, with_sign=with_sign, want_y=full, codes_kind=L.CODES_BF16,
                               kind="log")
    else:
        return _device_quant(x, mode, fsr, bit_width, with_sign), None
    return (y if full else eng.placeholder_like(x)), tag


def LogQuant(fsr=7, bit_width=3, with_sign=True, lin_back=True, _emit_codes=True):
    """sign(x) * 2^clamp(round(log2|x|), fsr - 2^bit_width, fsr)   (log_lin_connect.py:9-40)."""
    class _LogQuant(TaggingFunction):
        @staticmethod
        def forward(ctx, input):
            if not _emit_codes:          # weight quantizers: fp32 values only
                return _device_quant(input, L.Q_LOG, fsr, bit_width, with_sign)
            y, tag = _device_quant_tagged(input, L.Q_LOG, fsr, bit_width, with_sign)
            TaggingFunction._leave(tag)
            return y

        @staticmethod
        def backward(ctx, grad_output):
            if lin_back:
                return grad_output.clone()
            return _device_quant(grad_output, L.Q_LOG, fsr, bit_width, True)
    return _LogQuant


def LinQuant(fsr=7, bit_width=3, with_sign=True, lin_back=True, _emit_codes=True):
    """sign(x) * clamp(round(|x|/step) * step, 0, 2^fsr), step = 2^(fsr - bit_width)   (log_lin_connect.py:42-80).
    Unlike the reference (whose `step` is a CPU tensor, :65) this works on the device."""
    class _LinQuant(TaggingFunction):
        @staticmethod
        def forward(ctx, input):
            if bit_width == 32:
                return input.view_as(input)
            if not _emit_codes:
                return _device_quant(input, L.Q_LIN, fsr, bit_width, with_sign)
            y, tag = _device_quant_tagged(input, L.Q_LIN, fsr, bit_width, with_sign)
            TaggingFunction._leave(tag)
            return y

        @staticmethod
        def backward(ctx, grad_output):
            if bit_width == 32 or lin_back:
                return grad_output.clone()
            # sign(g) * clamp(round(g/step)*step, 0, 2^fsr)  (:78)
            return torch.sign(grad_output) * _device_quant(grad_output, L.Q_LIN, fsr, bit_width, False)
    return _LinQuant


def nnQuant(dtype="lin", fsr=7, bit_width=3, with_sign=True, lin_back=True, _emit_codes=True):
    """Module with a lin/log quantizer inside (log_lin_connect.py:84-100).  As an activation quantizer it also leaves the
    low-bit operand for the next quantized layer (`_emit_codes`; the layers build their weight quantizer without it)."""
    if dtype == "lin":
        return front(LinQuant(fsr=fsr, bit_width=bit_width, with_sign=with_sign, lin_back=lin_back, _emit_codes=_emit_codes))
    elif dtype == "log":
        return front(LogQuant(fsr=fsr, bit_width=bit_width, with_sign=with_sign, lin_back=lin_back, _emit_codes=_emit_codes))
    raise RuntimeError("Only 'log' and 'lin' dtype are supported !")


def Quant(input, dtype="lin", fsr=7, bit_width=3, with_sign=True, lin_back=True):
    """Functional lin/log quantizer (log_lin_connect.py:103-118)."""
    if dtype == "lin":
        return LinQuant(fsr=fsr, bit_width=bit_width, with_sign=with_sign, lin_back=lin_back).apply(input)
    elif dtype == "log":
        return LogQuant(fsr=fsr, bit_width=bit_width, with_sign=with_sign, lin_back=lin_back).apply(input)
    raise RuntimeError("Only 'log' and 'lin' dtype are supported !")
