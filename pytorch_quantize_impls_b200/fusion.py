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
            y, tag = ops.quant_act(x, L.Q_SIGN, want_y=full, codes_kind=eng.int_codes_kind(x), want_bits=eng.want_sign_bits(x), kind="sign", pre=pre)
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


def _bn_affine(bn):
    with torch.no_grad():
        s = torch.rsqrt(bn.running_var.float() + bn.eps)
        if bn.weight is not None:
            s = s * bn.weight.float()
        t = -bn.running_mean.float() * s
        if bn.bias is not None:
            t = t + bn.bias.float()
    return s.contiguous(), t.contiguous()


class FusedLayerQuant(nn.Module):
    """quantized Linear / Conv2d -> [BatchNorm (eval)] -> [Hardtanh | ReLU | ReLU6] -> activation quantizer with the
    quantizer running INSIDE the layer's tcgen05 epilogue (include/qtb200.h QtRequant): under
    `code_only_activations()` with autograd off, the epilogue writes the next layer's low-bit operand and the fp32
    activation never reaches HBM (bytes_module -> bytes_chained).  The BatchNorm affine is folded into the
    epilogue's per-column scale / bias.  Everywhere else (training, autograd, drop-in mode where the caller wants the
    fp32 tensor, shapes the tensor-core kernels decline) it is exactly the plain composition of its children."""

    def __init__(self, layer, bn, act, quant, consumer=None):
        super().__init__()
        self.layer = layer
        # the tail on its own is the one-pass BatchNorm+clamp+quantizer kernel: used whenever the epilogue fusion does not apply
        self.tail = FusedBNActQuant(bn, act, quant) if (bn is not None or act is not None) else quant
        self._consumer_needs_i8 = _needs_8bit_lanes(consumer)
        self._spec = None

    @property
    def bn(self):
        return self.tail.bn if isinstance(self.tail, FusedBNActQuant) else None

    @property
    def act(self):
        return self.tail.act if isinstance(self.tail, FusedBNActQuant) else None

    @property
    def quant(self):
        return self.tail.quant if isinstance(self.tail, FusedBNActQuant) else self.tail

    def _compose(self, x):
        return self.tail(self.layer(x))

    def _make_spec(self):
        kind, arg = self.quant._qt_spec
        lo, hi = _clamp_range(self.act)
        bn = self.bn
        key = None
        if bn is not None:
            params = [t for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var) if t is not None]
            key = tuple((t.data_ptr(), t._version) for t in params)
        if self._spec is not None and self._spec[0] == key:
            return self._spec[1]
        mul = add = None
        if bn is not None:
            mul, add = _bn_affine(bn)
        if kind == "dorefa" and arg == 1:
            kind = "sign"
        mode = {"sign": L.Q_SIGN, "ternary": L.Q_TERNARY, "dorefa": L.Q_DOREFA, "xnor": L.Q_XNOR_ROW}[kind]
        spec = eng.RequantSpec(mode, kind, bit_width=arg if kind == "dorefa" else 0, lo=lo, hi=hi, col_mul=mul, col_add=add)
        spec.force_8bit = self._consumer_needs_i8
        self._spec = (key, spec)
        return spec

    def forward(self, x):
        bn = self.bn
        fusable = (eng._code_only[0] and not torch.is_grad_enabled() and (x.is_cuda or x.is_meta)
                   and x.dim() == (4 if self.layer._is_conv else 2)
                   and (bn is None or (not bn.training and bn.track_running_stats and bn.running_mean is not None)))
        if fusable:
            try:
                return self.layer._forward_requant(x, self._make_spec())
            except eng.RequantUnsupported:
                pass
        return self._compose(x)

    def extra_repr(self):
        return "fused epilogue requant"


class FusedLayerBN(nn.Module):
    """quantized Linear / Conv2d -> BatchNorm (eval): the BatchNorm affine is folded into the per-column scale / bias of the
    layer's epilogue (y*s + t = acc*(cs*s) + (b*s + t)), so the normalisation costs no pass over the activation.  Applies
    with autograd off and the BatchNorm in eval mode; otherwise the plain composition runs."""

    def __init__(self, layer, bn, act=None):
        super().__init__()
        self.layer, self.bn, self.act = layer, bn, act        # bn and / or act (Hardtanh / ReLU / ReLU6) may be None
        self._spec = None

    def _make_spec(self):
        bn = self.bn
        key = None
        if bn is not None:
            params = [t for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var) if t is not None]
            key = tuple((t.data_ptr(), t._version) for t in params)
        if self._spec is None or self._spec[0] != key:
            mul, add = _bn_affine(bn) if bn is not None else (None, None)
            lo, hi = _clamp_range(self.act)
            self._spec = (key, eng.RequantSpec(-1, None, lo=lo, hi=hi, col_mul=mul, col_add=add))
        return self._spec[1]

    def forward(self, x):
        bn = self.bn
        n_out = self.layer.out_channels if self.layer._is_conv else self.layer.out_features
        if (not torch.is_grad_enabled() and (x.is_cuda or x.is_meta)
                # BatchNorm1d over a 3-D input normalises dim 1, not the layer's output features: only the plain ranks fold
                and x.dim() == (4 if self.layer._is_conv else 2)
                and (bn is None or (not bn.training and bn.track_running_stats and bn.running_mean is not None
                                    and bn.num_features == n_out))):
            return self.layer._forward_affine(x, self._make_spec())
        y = self.layer(x)
        if bn is not None:
            y = bn(y)
        return self.act(y) if self.act is not None else y

    def extra_repr(self):
        return "BatchNorm / clamp folded into the epilogue"


def _needs_8bit_lanes(consumer):
    """True when the layer that will read the codes cannot take e2m1 operands (DoReFa k >= 3 weights, convs, unknown)."""
    from .layers.common import QuantLayerMixin
    if consumer is None or not isinstance(consumer, QuantLayerMixin):
        return True
    if consumer._is_conv:
        return True
    bw = getattr(consumer, "bit_width", None)
    return bw is not None and bw > 2


def _is_quant_layer(m):
    from .layers.common import QuantLayerMixin
    return isinstance(m, QuantLayerMixin) and isinstance(m, (nn.Linear, nn.Conv2d))


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
            # quantized layer -> [BN] -> [clamp] -> quantizer: the quantizer moves into the layer's epilogue
            if _is_quant_layer(mods[i]) and not (mods[i]._is_conv and mods[i].groups != 1):
                j = i + 1
                bn = act = None
                want_bn = nn.BatchNorm2d if mods[i]._is_conv else nn.BatchNorm1d
                if j < len(mods) and isinstance(mods[j], want_bn):
                    bn = mods[j]
                    j += 1
                if j < len(mods) and isinstance(mods[j], (nn.Hardtanh, nn.ReLU, nn.ReLU6)) and not _is_quantizer(mods[j]):
                    act = mods[j]
                    j += 1
                if j < len(mods) and _is_quantizer(mods[j]) and (mods[j]._qt_spec[0] != "xnor" or not mods[i]._is_conv):
                    consumer = mods[j + 1] if j + 1 < len(mods) else None
                    out.append(FusedLayerQuant(mods[i], bn, act, mods[j], consumer))
                    i = j + 1
                    continue
                if bn is not None or act is not None:
                    # layer -> [BatchNorm] -> [clamp] with no quantizer behind it (pool / residual add follows): fold the
                    # normalisation and the clamp into the layer's epilogue
                    out.append(FusedLayerBN(mods[i], bn, act))
                    i += 1 + (bn is not None) + (act is not None)
                    continue
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


class OperandPrefetch(nn.Module):
    """Inference wrapper: at the start of every forward the transient tensor-core operands of ALL eval-mode quantized layers
    of `net` (the weight expansions, 5-15 us each) are launched on a side stream, so that they run beside the first
    activation quantizer instead of in front of each contraction; every layer then waits for its own operand only.
    The operand kind expanded for a layer is the one its previous forward asked for (the first forward runs unchanged).
    Works eagerly and inside CUDA-graph capture (the side stream forks from and re-joins the calling stream)."""

    def __init__(self, net):
        super().__init__()
        self.net = net
        self._side = None

    def forward(self, x):
        from .layers.common import QuantLayerMixin
        if torch.is_grad_enabled() or not (x.is_cuda or x.is_meta):
            return self.net(x)
        packs = []
        for m in self.net.modules():
            if isinstance(m, QuantLayerMixin) and not m.training:
                pack = m._current_pack()
                if pack.packed is not None and pack._last_kind is not None and pack._prefetch is None:
                    packs.append(pack)
        if not packs:
            return self.net(x)
        dev = packs[0].packed.device
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        self._side.wait_event(fork)
        with torch.cuda.stream(self._side):
            for pack in packs:
                out, ld = ops._expand_weight(pack, pack._last_kind)
                done = torch.cuda.Event()
                done.record(self._side)
                pack._prefetch = (pack._last_kind, out, ld, done)
        try:
            return self.net(x)
        finally:
            cur.wait_stream(self._side)          # re-join (also covers operands nobody consumed)
            for pack in packs:
                pack._prefetch = None


def prefetch_operands(net):
    """Wrap `net` (typically after fuse_inference) in an OperandPrefetch."""
    return OperandPrefetch(net)
