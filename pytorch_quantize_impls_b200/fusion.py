"""Inference-time fusion of the inter-layer pattern every reference net repeats:

    quantized conv / linear  ->  [pool]  ->  BatchNorm (eval)  ->  Hardtanh / ReLU  ->  activation quantizer

(benchmark/BinaryNet/AlexNetBin.py:13-48, MLPBin.py:42-53, models/samples/AlexNet_Dorefa.py:46-84).  `fuse_inference`
rewrites `BatchNorm -> clamp activation -> quantizer` runs inside nn.Sequential containers into ONE device pass
(`qt_quant_act` with the fused pre-transform x' = clamp(x * s[c] + t[c], lo, hi)): the fp32 conv output is read once and
the next layer's low-bit operand is written, instead of three read+write passes over the fp32 activation.

The fused value is computed as fma(x, s, t) with s = gamma / sqrt(var + eps), t = beta - mean * s, i.e. with one rounding
where the unfused chain has three; an activation that sits exactly on a rounding boundary of the quantizer can therefore
land one code away from the unfused result (same effect as between any two fp32 implementations of the chain).
"""
import torch
from torch import nn

from . import _engine as eng
from . import _lib as L
from . import _ops as ops
from .functions.common import _f32
from .functions.dorefa_connect import _quantize_with_codes

_BN = (nn.BatchNorm1d, nn.BatchNorm2d)


def _clamp_range(act):
    if act is None or isinstance(act, nn.Identity):
        return None, None
    if isinstance(act, nn.Hardtanh):
        return float(act.min_val), float(act.max_val)
    if isinstance(act, nn.ReLU6):
        return 0.0, 6.0
    if isinstance(act, nn.ReLU):
        return 0.0, float("inf")
    return NotImplemented, NotImplemented


def _is_quantizer(m):
    return getattr(m, "_qt_spec", None) is not None


class FusedBNActQuant(nn.Module):
    """BatchNorm(eval) -> clamp -> activation quantizer in one kernel launch.  Falls back to the plain composition
    whenever the fused form does not apply (training mode, autograd on, BatchNorm using batch statistics)."""

    def __init__(self, bn, act, quant):
        super().__init__()
        self.bn, self.act, self.quant = bn, act, quant
        self._cache = None

    def _compose(self, x):
        if self.bn is not None:
            x = self.bn(x)
        if self.act is not None:
            x = self.act(x)
        return self.quant(x)

    def _affine(self, x):
        bn = self.bn
        C = x.shape[1]
        if bn is None:
            key = ("none", C, x.device)
            if self._cache is None or self._cache[0] != key:
                self._cache = (key, torch.ones(C, device=x.device), torch.zeros(C, device=x.device))
            return self._cache[1], self._cache[2]
        params = [t for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var) if t is not None]
        key = tuple((t.data_ptr(), t._version) for t in params)
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                s = torch.rsqrt(bn.running_var.float() + bn.eps)
                if bn.weight is not None:
                    s = s * bn.weight.float()
                t = -bn.running_mean.float() * s
                if bn.bias is not None:
                    t = t + bn.bias.float()
            self._cache = (key, s.contiguous(), t.contiguous())
        return self._cache[1], self._cache[2]

    def forward(self, x):
        bn = self.bn
        fusable = (not torch.is_grad_enabled() and x.is_cuda and x.dim() in (2, 4)
                   and (bn is None or (not bn.training and bn.track_running_stats and bn.running_mean is not None)))
        if not fusable:
            return self._compose(x)
        lo, hi = _clamp_range(self.act)
        s, t = self._affine(x)
        pre = (s, t, lo, hi)
        kind, arg = self.quant._qt_spec
        full = eng.want_fp32_result(x)
        if kind == "dorefa":
            y, tag = _quantize_with_codes(x, arg, pre=pre)
        elif kind == "sign":
            y, tag = ops.quant_act(x, L.Q_SIGN, want_y=full, codes_kind=eng.int_codes_kind(x), want_bits=(x.dim() == 2), kind="sign", pre=pre)
            y = y if full else eng.placeholder_like(x)
        elif kind == "ternary":
            y, tag = ops.quant_act(x, L.Q_TERNARY, want_y=full, codes_kind=eng.int_codes_kind(x), kind="ternary", pre=pre)
            y = y if full else eng.placeholder_like(x)
        elif kind == "xnor" and x.dim() == 2:
            y, tag = ops.quant_act(x, L.Q_XNOR_ROW, want_y=full, codes_kind=eng.xnor_codes_kind(), want_row_scale=True,
                                   kind="xnor", pre=pre)
            y = y if full else eng.placeholder_like(x)
        else:
            return self._compose(x)
        return eng.attach_tag(y, tag)

    def extra_repr(self):
        return "fused"


def fuse_inference(module):
    """Rewrite, in place and recursively, every `[BatchNorm] -> [Hardtanh|ReLU|ReLU6] -> activation quantizer` run found
    inside nn.Sequential containers into a FusedBNActQuant.  Call it after loading weights (the fused module keeps
    the original sub-modules as children `bn`, `act`, `quant`).  Returns `module`."""
    for name, child in list(module.named_children()):
        fuse_inference(child)
    if isinstance(module, nn.Sequential):
        mods = list(module.children())
        out, i = [], 0
        while i < len(mods):
            j = i
            bn = act = None
            if isinstance(mods[j], _BN):
                bn = mods[j]
                j += 1
            if j < len(mods) and _clamp_range(mods[j])[0] is not NotImplemented and not _is_quantizer(mods[j]) \
                    and isinstance(mods[j], (nn.Hardtanh, nn.ReLU, nn.ReLU6)):
                act = mods[j]
                j += 1
            if j < len(mods) and _is_quantizer(mods[j]) and (bn is not None or act is not None):
                out.append(FusedBNActQuant(bn, act, mods[j]))
                i = j + 1
            else:
                out.append(mods[i])
                i += 1
        if len(out) != len(mods):
            for k in list(module._modules.keys()):
                del module._modules[k]
            for k, m in enumerate(out):
                module.add_module(str(k), m)
    return module
