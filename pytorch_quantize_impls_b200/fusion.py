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
        # a conv consumer reads the codes through TMA im2col: give it whole 128-byte channel rows (engine.conv_pad_channels)
        self._pad_channels = 0
        if layer._is_conv and _is_quant_layer(consumer) and consumer._is_conv and consumer.groups == 1:
            self._pad_channels = eng.conv_pad_channels(layer.out_channels)
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
        spec.pad_channels = self._pad_channels
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


def _pool_params(pool):
    """(kernel, stride, padding) pairs of an nn.MaxPool2d the pooling kernels can run, else None."""
    if not isinstance(pool, nn.MaxPool2d) or pool.ceil_mode or pool.return_indices:
        return None
    two = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    if two(pool.dilation) != (1, 1):
        return None
    k = two(pool.kernel_size)
    s = two(pool.stride if pool.stride is not None else pool.kernel_size)
    p = two(pool.padding)
    if 2 * p[0] > k[0] or 2 * p[1] > k[1]:
        return None
    return k, s, p


def _bn_ready(bn):
    return bn is None or (not bn.training and bn.track_running_stats and bn.running_mean is not None)


class FusedLayerPoolQuant(FusedLayerQuant):
    """quantized Conv2d -> MaxPool2d -> [BatchNorm (eval)] -> [clamp] -> activation quantizer (models/Alexnet/Alexnet_Bin.py:13-22,
    benchmark/BinaryNet/AlexNetBin.py:13-24): the conv's tcgen05 epilogue applies BatchNorm + clamp + quantizer and writes
    channels-last 8-bit codes, and the pool runs on the codes (qt_pool_codes: 1 byte per element instead of 4 + 4).  An activation
    quantizer is monotone, so pooling commutes with it; where the folded BatchNorm scale of a channel is negative the pool of that
    channel is a min-pool.  Falls back to the plain composition wherever FusedLayerQuant would."""

    def __init__(self, layer, pool, bn, act, quant, consumer=None):
        super().__init__(layer, bn, act, quant, consumer)
        self.pool = pool
        self._mask = None

    def _compose(self, x):
        return self.tail(self.pool(self.layer(x)))

    def _min_mask(self, spec):
        if spec.col_mul is None:
            return None
        key = (spec.col_mul.data_ptr(), spec.col_mul._version)
        if self._mask is None or self._mask[0] != key:
            neg = spec.col_mul < 0
            self._mask = (key, neg.to(torch.uint8).contiguous() if bool(neg.any()) else None)
        return self._mask[1]

    def forward(self, x):
        geo = _pool_params(self.pool)
        fusable = (geo is not None and eng._code_only[0] and not torch.is_grad_enabled() and (x.is_cuda or x.is_meta)
                   and x.dim() == 4 and self.layer.out_channels % 16 == 0 and _bn_ready(self.bn))
        if fusable:
            spec = self._make_spec()
            try:
                y = self.layer._forward_requant(x, spec)
            except eng.RequantUnsupported:
                return self._compose(x)
            tag = eng.get_tag(y)
            pooled = ops.pool_codes(tag.codes, *geo, use_min=self._min_mask(spec))
            B, PH, PW, Cn = pooled.shape                    # Cn: channel pitch of the codes (>= out_channels when padded)
            tag.codes, tag.rows, tag.cols, tag.ld = pooled, B, Cn * PH * PW, Cn * PH * PW
            out = torch.empty((B, self.layer.out_channels, PH, PW), dtype=torch.float32, device="meta")
            return eng.attach_tag(out, tag)
        return self._compose(x)

    def extra_repr(self):
        return "fused epilogue requant + pool on codes"


class FusedConvPool(nn.Module):
    """quantized Conv2d -> [BatchNorm (eval)] -> [clamp] -> MaxPool2d with an fp32 result (the stem of a residual net,
    models/Resnet/Resnet_bin.py:68-69 + pool): BatchNorm and clamp fold into the conv epilogue, which writes a channels-last
    fp32 tensor with full-line TMA stores; the pool reads it once (qt_pool_quant_f32) and returns a channels-last tensor."""

    def __init__(self, layer, bn, act, pool):
        super().__init__()
        self.inner = FusedLayerBN(layer, bn, act)
        self.pool = pool
        self._next = None          # _qt_spec of the DoReFa quantizer that reads the pooled tensor next (a residual block's q_in)

    def forward(self, x):
        geo = _pool_params(self.pool)
        lay = self.inner.layer
        if (geo is not None and not torch.is_grad_enabled() and x.is_cuda and x.dim() == 4 and lay._is_conv
                and lay.out_channels % 4 == 0 and _bn_ready(self.inner.bn)):
            y = lay._forward_affine(x, self.inner._make_spec(), out_format="nhwc")
            nxt = self._next
            lo, hi = _clamp_range(self.inner.act)
            if (nxt is not None and nxt[0] == "dorefa" and 2 <= nxt[1] <= 8 and eng._code_only[0] and lay.out_channels % 16 == 0):
                # the pool pass also writes the next quantizer's codes: one read of the conv output, no separate quantizer pass
                k = nxt[1]
                ck = L.CODES_U8 if k == 8 else L.CODES_I8
                out, codes, ovf = ops.pool_quant_f32(y, *geo, want_out=True, mode=L.Q_DOREFA, bit_width=k, codes_kind=ck)
                tag = ops.ActCodes()
                B, PH, PW, Cn = codes.shape
                tag.kind, tag.bit_width, tag.codes, tag.codes_kind = "dorefa", k, codes, ck
                tag.rows, tag.cols, tag.ld, tag.layout = B, Cn * PH * PW, Cn * PH * PW, "nhwc"
                tag.scale = _f32(_f32(1.0) / _f32(2 ** k - 1))
                tag.row_sum = tag.row_scale = tag.bits = None
                tag.ld_bits, tag.row_parts, tag.row_mul, tag.overflow = 0, 0, 1.0, ovf
                tag.range_ok = ops.clamp_guarantees_lane(ck, k, lo, hi) if lo is not NotImplemented else False
                return eng.attach_tag(out, tag)
            return ops.pool_quant_f32(y, *geo)[0]
        return self.pool(self.inner(x))


class FlattenCodes(nn.Module):
    """Flatten between the conv stack and the classifier of a fused chain.  The conv chain leaves channels-last codes
    [B, H, W, C]; their flatten is (h, w, c)-ordered, so the Linear that follows is told to permute its weight columns once
    (QuantLayerMixin._in_perm) and then reads the codes as they are -- no transpose pass, no fp32 detour."""

    def __init__(self, flatten, consumer):
        super().__init__()
        self.flatten = flatten
        self._consumer = [consumer]            # not a registered child: it already lives in the parent container

    def forward(self, x):
        cons = self._consumer[0]
        if x.dim() != 4 or torch.is_grad_enabled() or not (x.is_cuda or x.is_meta):
            if cons._in_perm is not None:
                raise RuntimeError("this network was re-ordered for inference by fuse_inference (channels-last flatten): run it under "
                                   "torch.no_grad() on 4-D CUDA inputs")
            return self.flatten(x)
        B, Cn, H, W = x.shape
        if cons._in_perm is None:
            cons._set_in_perm((Cn, H, W))
        elif cons._in_perm != (Cn, H, W):
            raise RuntimeError("FlattenCodes: activation shape changed after the classifier weights were re-ordered")
        tag = eng.get_tag(x)
        if x.is_meta:
            out = torch.empty((B, Cn * H * W), dtype=torch.float32, device="meta")
        else:
            out = x.permute(0, 2, 3, 1).reshape(B, -1)
        if tag is not None and tag.layout == "nhwc" and tag.codes.dim() == 4 and (Cn * H * W) % 16 == 0:
            t2 = ops.ActCodes()
            for f in ops.ActCodes.__slots__:
                if hasattr(tag, f):
                    setattr(t2, f, getattr(tag, f))
            t2.codes = tag.codes.view(B, H * W * Cn)
            t2.rows, t2.cols, t2.ld, t2.layout = B, H * W * Cn, H * W * Cn, "rows"
            t2.row_sum, t2.row_parts, t2.row_mul = None, 0, 1.0
            if getattr(cons, "bit_width", None) == 8 and tag.kind == "dorefa":
                t2.row_sum = ops.rowsum_codes(t2.codes)         # zero point of unsigned 8-bit weights
            return eng.attach_tag(out, t2)
        if x.is_meta:
            raise RuntimeError("FlattenCodes: the code-only activation cannot be flattened for this classifier")
        return out


class FusedActLayer(nn.Module):
    """activation quantizer -> quantized Linear at the head of a chain, where the input is a real fp32 tensor.  With
    `set_banded_head(True)` (off by default: slower than the plain pair on B200, see engine.set_banded_head) and under
    `code_only_activations()` the pair runs as a two-stream pipeline over row bands (engine.linear_banded) -- the HBM-bound
    quantizer of band i+1 beside the tensor-bound contraction of band i -- instead of one after the other."""
    MIN_ROWS = 4096

    def __init__(self, quant, inner):
        super().__init__()
        self.quant, self.inner = quant, inner

    def _layer(self):
        return self.inner.layer if isinstance(self.inner, FusedLayerBN) else self.inner

    def forward(self, x):
        lay = self._layer()
        kind, arg = self.quant._qt_spec
        ok = ((eng._banded_head[0] or eng._overlap_head[0]) and eng._code_only[0] and not torch.is_grad_enabled() and x.is_cuda
              and x.dim() == 2
              and x.dtype == torch.float32
              and x.shape[0] >= self.MIN_ROWS and x.shape[1] % 1024 == 0 and x.is_contiguous() and not lay.training
              and not (isinstance(self.inner, FusedLayerBN) and not _bn_ready(self.inner.bn)))
        if not ok:
            return self.inner(self.quant(x))
        if kind == "dorefa" and arg == 1:
            kind = "sign"
        pack = lay._current_pack()
        int_w = pack.kind in ("sign", "ternary", "dorefa", "lin")
        if kind == "xnor":
            if pack.kind not in ("xnor", "sign", "ternary", "dorefa"):
                return self.inner(self.quant(x))
        elif not int_w or not (kind in ("sign", "ternary") or 2 <= arg <= 8):
            return self.inner(self.quant(x))
        # 1-bit / ternary / 2-bit codes ride the e2m1 lane when the weights can meet them there, else an 8-bit lane
        f4 = eng._fp4[0] and eng._f4_weight_ok(pack) and (kind in ("sign", "ternary") or arg == 2)
        small = L.CODES_F4 if f4 else L.CODES_I8

        def quantize(rows, max_ctas, ready=None, ready_rows=0, codes_out=None):
            kw = dict(want_y=False, max_ctas=max_ctas, ready=ready, ready_rows=ready_rows, codes_out=codes_out)
            if kind == "sign":
                return ops.quant_act(rows, L.Q_SIGN, codes_kind=small, kind="sign", **kw)[1]
            if kind == "ternary":
                return ops.quant_act(rows, L.Q_TERNARY, codes_kind=small, kind="ternary", **kw)[1]
            if kind == "xnor":
                return ops.quant_act(rows, L.Q_XNOR_ROW, codes_kind=eng.xnor_codes_kind(), want_row_scale=True, kind="xnor", **kw)[1]
            ck = small if arg == 2 else (L.CODES_I8 if arg <= 7 else L.CODES_U8)
            tag = ops.quant_act(rows, L.Q_DOREFA, bit_width=arg, codes_kind=ck, want_row_sum=True, kind="dorefa", **kw)[1]
            tag.scale = _f32(_f32(1.0) / _f32(2 ** arg - 1))
            return tag

        affine = self.inner._make_spec() if isinstance(self.inner, FusedLayerBN) else None
        if (eng._overlap_head[0] and not eng._banded_head[0] and x.shape[0] % eng.OVERLAP_ROWS == 0
                and kind in ("sign", "ternary")):
            return eng.linear_overlapped(x, quantize, small, pack, lay.bias, affine=affine)
        if not eng._banded_head[0]:
            return self.inner(self.quant(x))
        return eng.linear_banded(x, quantize, pack, lay.bias, affine=affine)

    def extra_repr(self):
        return "quantizer beside the contraction (progress counters) / banded pipeline"


class FusedBasicBlock(nn.Module):
    """Residual block (models/Resnet/Resnet_bin.py:7-32 with k-bit activations, nets.TerBasicBlock) for inference chains:

        xq   = q_in(x)                                   codes (left by the previous block's epilogue, else one quantizer pass)
        c1   = conv1(xq) -> BN -> clamp -> quantizer     requant epilogue, channels-last codes
        res  = x  |  shortcut conv(xq) -> BN             channels-last fp32
        out  = clamp(BN(conv2(c1)) + res)                ONE epilogue: residual add + clamp + fp32 channels-last store (TMA)
                                                         + the NEXT block's q_in codes

    so the residual stream is read once and written once per block (plus 1 byte of codes) instead of the add / clamp / quantizer
    passes.  Outside code-only inference it is the wrapped block."""

    def __init__(self, block, next_spec=None, next_reads_fp32=True):
        super().__init__()
        self.block = block
        self._next = next_spec                 # _qt_spec of the next block's input quantizer (None: no codes needed)
        # False: the next block has a shortcut conv (it reads this block's output through its input quantizer only), so the
        # fp32 block output is not written at all -- only the codes
        self._next_reads_fp32 = next_reads_fp32
        self._spec2 = None

    def _parts(self):
        b = self.block
        f1 = b.branch1[0] if len(b.branch1) == 1 else None
        f2 = b.branch2[0] if len(b.branch2) == 1 else None
        fs = None if b.shortcut is None else (b.shortcut[0] if len(b.shortcut) == 1 else NotImplemented)
        if not isinstance(f1, FusedLayerQuant) or not isinstance(f2, FusedLayerBN) or fs is NotImplemented \
                or (fs is not None and not isinstance(fs, FusedLayerBN)):
            return None
        return f1, f2, fs

    def _make_spec2(self, f2, lo, hi):
        bn = f2.bn
        key = None
        if bn is not None:
            key = tuple((t.data_ptr(), t._version) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var) if t is not None)
        if self._spec2 is None or self._spec2[0] != key:
            mul, add = _bn_affine(bn) if bn is not None else (None, None)
            plain = eng.RequantSpec(-1, None, lo=lo, hi=hi, col_mul=mul, col_add=add)
            rq = None
            if self._next is not None:
                rq = eng.RequantSpec(L.Q_DOREFA, "dorefa", bit_width=self._next[1], lo=lo, hi=hi, col_mul=mul, col_add=add)
                rq.force_8bit = True
            self._spec2 = (key, plain, rq)
        return self._spec2[1], self._spec2[2]

    def forward(self, x):
        parts = self._parts()
        b = self.block
        spec_in = getattr(b.q_in, "_qt_spec", None)
        lo, hi = _clamp_range(b.clip)
        tag = eng.get_tag(x)
        # a code-only input (no fp32 tensor) is enough when the shortcut is a conv on the codes
        codes_only_in = bool(x.is_meta and tag is not None and parts is not None and parts[2] is not None)
        ok = (parts is not None and eng._code_only[0] and not torch.is_grad_enabled() and x.dim() == 4
              and (x.is_cuda or codes_only_in)
              and x.dtype == torch.float32 and spec_in is not None and spec_in[0] == "dorefa" and 2 <= spec_in[1] <= 8
              and lo is not NotImplemented and x.shape[1] % 16 == 0
              and (self._next is None or (self._next[0] == "dorefa" and 2 <= self._next[1] <= 8)))
        if ok:
            f1, f2, fs = parts
            ok = _bn_ready(f1.bn) and _bn_ready(f2.bn) and (fs is None or _bn_ready(fs.bn)) and f2.act is None \
                and f2.layer.out_channels % 32 == 0
        abits = spec_in[1] if ok else 0
        tag_ok = tag is not None and tag.layout == "nhwc" and tag.kind == "dorefa" and tag.bit_width == abits
        if ok and x.is_meta and not tag_ok:
            ok = False
        if not ok:
            if x.is_meta:
                raise RuntimeError("FusedBasicBlock: received a code-only activation it cannot consume (the producer block was "
                                   "told that this block reads codes only)")
            return b(x)
        if not tag_ok:
            if not ops.is_channels_last(x):
                x = x.contiguous(memory_format=torch.channels_last)
            tag = ops.quant_act_nhwc(x, L.Q_DOREFA, bit_width=abits, codes_kind=L.CODES_U8 if abits == 8 else L.CODES_I8,
                                     kind="dorefa")
            tag.scale = _f32(_f32(1.0) / _f32(2 ** abits - 1))
        elif not ops.is_channels_last(x):
            x = x.contiguous(memory_format=torch.channels_last)
        xq = eng.attach_tag(torch.empty(x.shape, dtype=torch.float32, device="meta"), tag)
        c1 = f1(xq)
        if eng.get_tag(c1) is None or not c1.is_meta:
            if x.is_meta:
                raise RuntimeError("FusedBasicBlock: the first conv declined the fused epilogue on a code-only input")
            return b(x)                        # the first conv declined the fused epilogue: plain block
        res = x if fs is None else fs.layer._forward_affine(xq, fs._make_spec(), out_format="nhwc")
        plain, rq = self._make_spec2(f2, lo, hi)
        if rq is not None:
            try:
                return f2.layer._forward_requant(c1, rq, out_format="nhwc", residual=res, keep_out=self._next_reads_fp32)
            except eng.RequantUnsupported:
                pass
        return f2.layer._forward_affine(c1, plain, out_format="nhwc", residual=res)

    def extra_repr(self):
        return "residual add + clamp + next quantizer in the second conv's epilogue"


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


def _is_flatten(m):
    return isinstance(m, nn.Flatten) and m.start_dim == 1 and m.end_dim == -1 or type(m).__name__ == "Flatten" and not list(m.parameters())


def _codes_only_consumer(blk):
    """The block can run from its input CODES alone: fusable parts (checked again at run time) and a conv shortcut."""
    if blk.shortcut is None or len(blk.shortcut) != 1 or not isinstance(blk.shortcut[0], FusedLayerBN):
        return False
    return (len(blk.branch1) == 1 and isinstance(blk.branch1[0], FusedLayerQuant)
            and len(blk.branch2) == 1 and isinstance(blk.branch2[0], FusedLayerBN))


def _is_block(m):
    """Residual block with the attribute layout of nets.TerBasicBlock."""
    return all(hasattr(m, a) for a in ("q_in", "branch1", "branch2", "clip", "shortcut")) and not isinstance(m, FusedBasicBlock)


class FusedAvgLinear:
    """AdaptiveAvgPool2d(1) -> flatten -> nn.Linear (fp32 head of the residual nets) as ONE batch-invariant kernel (qt_head_f32)
    on the inference path; anything else (autograd, CPU, other dtypes) runs the two modules.  Not an nn.Module: it only
    borrows the parameters of the modules it was built from."""

    def __init__(self, avg, linear):
        self.avg, self.linear = avg, linear

    def __call__(self, x):
        lin = self.linear
        if (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and not torch.is_grad_enabled()
                and lin.weight.dtype == torch.float32 and lin.weight.is_cuda):
            return ops.head_f32(x, lin.weight, lin.bias)
        return lin(self.avg(x).flatten(1))


def fuse_inference(module):
    """Rewrite, in place and recursively, the inter-layer patterns of the reference nets found inside nn.Sequential containers:

        layer -> [BatchNorm] -> [clamp] -> quantizer                 FusedLayerQuant      (quantizer in the layer's epilogue)
        conv  -> MaxPool -> [BatchNorm] -> [clamp] -> quantizer      FusedLayerPoolQuant  (+ pool on the codes)
        layer -> [BatchNorm] -> [clamp]                              FusedLayerBN         (affine + clamp in the epilogue)
        conv  -> [BatchNorm] -> [clamp] -> MaxPool                   FusedConvPool        (channels-last fp32 + one pool pass)
        [BatchNorm] -> [clamp] -> quantizer                          FusedBNActQuant      (one elementwise pass)
        Flatten between a conv chain and a quantized Linear          FlattenCodes         (channels-last flatten, weights re-ordered)
        quantizer -> quantized Linear at the head of a chain         FusedActLayer        (banded quantizer / contraction pipeline)
        residual blocks (nets.TerBasicBlock)                         FusedBasicBlock      (add + clamp + next quantizer in conv2)

    Call it after loading weights (the fused modules keep the original sub-modules as children).  Returns `module`."""
    for name, child in list(module.named_children()):
        fuse_inference(child)
    if not isinstance(module, nn.Sequential):
        # residual nets (nets.ResNetTer layout): the stem's pool pass also emits the codes of the first block's input quantizer
        stem, layers = getattr(module, "stem", None), getattr(module, "layers", None)
        if (isinstance(stem, nn.Sequential) and isinstance(layers, nn.Sequential) and len(stem) and len(layers)
                and isinstance(stem[-1], FusedConvPool) and isinstance(layers[0], FusedBasicBlock)):
            stem[-1]._next = getattr(layers[0].block.q_in, "_qt_spec", None)
        avg, lin = getattr(module, "avg", None), getattr(module, "linear", None)
        if (getattr(type(module), "_uses_fused_head", False) and isinstance(avg, nn.AdaptiveAvgPool2d)
                and avg.output_size in (1, (1, 1)) and type(lin) is nn.Linear):
            module.__dict__["_fused_head"] = FusedAvgLinear(avg, lin)
        return module
    mods = list(module.children())
    out, i = [], 0
    clamp_t = (nn.Hardtanh, nn.ReLU, nn.ReLU6)
    while i < len(mods):
        m = mods[i]
        if _is_quant_layer(m) and not (m._is_conv and m.groups != 1):
            j = i + 1
            pool = bn = act = None
            want_bn = nn.BatchNorm2d if m._is_conv else nn.BatchNorm1d
            if m._is_conv and j < len(mods) and _pool_params(mods[j]) is not None:
                pool = mods[j]
                j += 1
            if j < len(mods) and isinstance(mods[j], want_bn):
                bn = mods[j]
                j += 1
            if j < len(mods) and isinstance(mods[j], clamp_t) and not _is_quantizer(mods[j]):
                act = mods[j]
                j += 1
            if j < len(mods) and _is_quantizer(mods[j]) and (mods[j]._qt_spec[0] != "xnor" or not m._is_conv):
                # the quantizer moves into the layer's epilogue (a pool in between runs on the codes)
                k = j + 1
                consumer = mods[k] if k < len(mods) else None
                if consumer is not None and _is_flatten(consumer) and k + 1 < len(mods):
                    consumer = mods[k + 1]
                if pool is not None:
                    out.append(FusedLayerPoolQuant(m, pool, bn, act, mods[j], consumer))
                else:
                    out.append(FusedLayerQuant(m, bn, act, mods[j], consumer))
                i = j + 1
                continue
            if pool is None and (bn is not None or act is not None):
                if m._is_conv and j < len(mods) and _pool_params(mods[j]) is not None:
                    out.append(FusedConvPool(m, bn, act, mods[j]))       # conv -> BN -> clamp -> pool (fp32 result)
                    i = j + 1
                    continue
                # layer -> [BatchNorm] -> [clamp] with no quantizer behind it (pool / residual add follows)
                out.append(FusedLayerBN(m, bn, act))
                i = j
                continue
        j = i
        bn = act = None
        if isinstance(mods[j], _BN):
            bn = mods[j]
            j += 1
        if j < len(mods) and isinstance(mods[j], clamp_t) and not _is_quantizer(mods[j]):
            act = mods[j]
            j += 1
        if j < len(mods) and _is_quantizer(mods[j]) and (bn is not None or act is not None):
            out.append(FusedBNActQuant(bn, act, mods[j]))
            i = j + 1
        else:
            out.append(mods[i])
            i += 1
    # second pass over the rewritten list: flatten on codes, banded head pair, residual blocks
    res, i = [], 0
    while i < len(out):
        m = out[i]
        nxt = out[i + 1] if i + 1 < len(out) else None
        if (_is_flatten(m) and res and isinstance(res[-1], (FusedLayerQuant, FusedBNActQuant)) and nxt is not None
                and _is_quant_layer(_unwrap(nxt)) and not _unwrap(nxt)._is_conv and hasattr(_unwrap(nxt), "_set_in_perm")
                and type(_unwrap(nxt)).__name__ != "LinearXNOR"):
            src = res[-1]
            if isinstance(src, FusedLayerQuant) and src.layer._is_conv:
                res.append(FlattenCodes(m, _unwrap(nxt)))
                i += 1
                continue
        if (_is_quantizer(m) and not isinstance(m, (FusedBNActQuant,)) and nxt is not None
                and (_is_quant_layer(nxt) or isinstance(nxt, FusedLayerBN)) and not _unwrap(nxt)._is_conv
                and not (res and isinstance(res[-1], FlattenCodes))):
            res.append(FusedActLayer(m, nxt))
            i += 2
            continue
        if _is_block(m):
            nspec = getattr(nxt.q_in, "_qt_spec", None) if (nxt is not None and _is_block(nxt)) else None
            # the next block adds `shortcut(q_in(x))` instead of x itself: it never reads this block's fp32 output
            reads_fp32 = not (nspec is not None and nxt.shortcut is not None and _codes_only_consumer(nxt))
            res.append(FusedBasicBlock(m, nspec, reads_fp32))
            i += 1
            continue
        res.append(m)
        i += 1
    if len(res) != len(mods) or any(a is not b for a, b in zip(res, mods)):
        for k in list(module._modules.keys()):
            del module._modules[k]
        for k, m in enumerate(res):
            module.add_module(str(k), m)
    return module


def _unwrap(m):
    return m.layer if isinstance(m, (FusedLayerQuant, FusedLayerBN)) else m


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
