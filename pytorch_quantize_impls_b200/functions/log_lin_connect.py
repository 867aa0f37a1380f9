"""Miyashita lin/log fixed-point quantizers -- surface of QuantTorch/functions/log_lin_connect.py."""
import torch

from .. import _lib as L
from .. import _ops as ops
from .common import front


def _device_quant(x, mode, fsr, bit_width, with_sign):
    return ops.quant_act(x, mode, bit_width=bit_width, fsr=fsr, with_sign=with_sign, want_y=True)[0]


def LogQuant(fsr=7, bit_width=3, with_sign=True, lin_back=True):
    """sign(x) * 2^clamp(round(log2|x|), fsr - 2^bit_width, fsr)   (log_lin_connect.py:9-40)."""
    class _LogQuant(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input):
            return _device_quant(input, L.Q_LOG, fsr, bit_width, with_sign)

        @staticmethod
        def backward(ctx, grad_output):
            if lin_back:
                return grad_output.clone()
            return _device_quant(grad_output, L.Q_LOG, fsr, bit_width, True)
    return _LogQuant


def LinQuant(fsr=7, bit_width=3, with_sign=True, lin_back=True):
    """sign(x) * clamp(round(|x|/step) * step, 0, 2^fsr), step = 2^(fsr - bit_width)   (log_lin_connect.py:42-80).
    Unlike the reference (whose `step` is a CPU tensor, :65) this works on the device."""
    class _LinQuant(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input):
            if bit_width == 32:
                return input.view_as(input)
            return _device_quant(input, L.Q_LIN, fsr, bit_width, with_sign)

        @staticmethod
        def backward(ctx, grad_output):
            if bit_width == 32 or lin_back:
                return grad_output.clone()
            # sign(g) * clamp(round(g/step)*step, 0, 2^fsr)  (:78)
            return torch.sign(grad_output) * _device_quant(grad_output, L.Q_LIN, fsr, bit_width, False)
    return _LinQuant


def nnQuant(dtype="lin", fsr=7, bit_width=3, with_sign=True, lin_back=True):
    """Module with a lin/log quantizer inside (log_lin_connect.py:84-100)."""
    if dtype == "lin":
        return front(LinQuant(fsr=fsr, bit_width=bit_width, with_sign=with_sign, lin_back=lin_back))
    elif dtype == "log":
        return front(LogQuant(fsr=fsr, bit_width=bit_width, with_sign=with_sign, lin_back=lin_back))
    raise RuntimeError("Only 'log' and 'lin' dtype are supported !")


def Quant(input, dtype="lin", fsr=7, bit_width=3, with_sign=True, lin_back=True):
    """Functional lin/log quantizer (log_lin_connect.py:103-118)."""
    if dtype == "lin":
        return LinQuant(fsr=fsr, bit_width=bit_width, with_sign=with_sign, lin_back=lin_back).apply(input)
    elif dtype == "log":
        return LogQuant(fsr=fsr, bit_width=bit_width, with_sign=with_sign, lin_back=lin_back).apply(input)
    raise RuntimeError("Only 'log' and 'lin' dtype are supported !")
