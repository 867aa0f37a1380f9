"""Contraction routing: (activation operand kind) x (weight pack kind) -> kernel + epilogue.

                        | W sign / ternary / DoReFa / Lin (integer codes)     | W xnor (alpha[k] * sign), log (2^e), real planes
  ----------------------+------------------------------------------------------+--------------------------------------------------
  A e2m1 codes          | tcgen05 kind::mxf4, unit block scales (sign, ternary, | re-read as real values (row below); not available
  (sign/ter/DoReFa-2,   | DoReFa <= 2 weights); other integer weights: codes    | in code-only mode
   2-D inputs)          | re-emitted as int8                                    |
  A int8/uint8 codes    | tcgen05 kind::i8 (XNOR-popcount kernels for 1-bit x   | same
  (DoReFa-k, Lin, conv) | {1-bit, ternary} when M <= 64); conv: TMA im2col      |
  A fp16 codes (xnor)   | kind::f16 on exact fp16 codes, row_scale = row mean   | xnor: ONE kind::f16 pass on fp16(alpha[k]/alpha_max * s),
                        |                                                       | alpha_max in the epilogue (bf16x2 mode: 2 passes)
  A bf16 values (log)   | kind::f16 (bf16), 1 pass                              | log: 1 pass; xnor / real: 2 passes (W hi, W lo)
  A real (untagged)     | bf16 hi/mid/lo split of A, 3 passes (conv first       | 5 passes (3 A planes x W hi, 2 x W lo); log: 3
                        | layers: hi/lo gather+split, 2 passes)                 |

Integer accumulators are exact (kind::i8: s32; kind::mxf4: integers in fp32 for K < 2^22); the real-activation route carries 24
(16 for first conv layers) significant bits of the activation, real-valued weights 16 bits, far inside the 1e-3 tolerance of
the north star.  Epilogue options shared by every tcgen05 route: scale / row / column scales, bias, folded BatchNorm, output
clamp, NCHW addressing, raw accumulators, and the fused re-quantisation that writes the next layer's operand (RequantSpec).
Large tiles run on CTA pairs (cta_group::2); everything a tensor route declines falls to the CUDA-core kernels.
"""
import os
import threading

import torch

from . import _lib as L
from . import _ops as ops

# M at or below which the 1-bit CUDA-core XNOR+popcount kernels are preferred over expanding the weights
# for the tensor cores (GEMV-like shapes: the packed weights are read once, 1 bit each).
POPCOUNT_MAX_M = int(os.environ.get("QTB200_POPCOUNT_MAX_M", "64"))
_force_backend = {"i8": L.BACKEND_AUTO, "bf16": L.BACKEND_AUTO}
_force_popcount = [False]
_implicit_conv = [os.environ.get("QTB200_IMPLICIT_CONV", "1") != "0"]
# XnorNet product precision: "fp16" = one fp16 tensor pass (alpha[k]*sign rounded to 11 bits, ~1e-4 of max|y|),
# "bf16x2" = two bf16 passes over hi/lo weight planes (~1e-5).  Both are inside the 1e-3 tolerance.
_xnor_mode = ["fp16"]


def set_xnor_mode(mode):
    if mode not in ("fp16", "bf16x2"):
        raise ValueError("xnor mode must be 'fp16' or 'bf16x2'")
    _xnor_mode[0] = mode


def xnor_codes_kind():
    return L.CODES_F16 if _xnor_mode[0] == "fp16" else L.CODES_BF16


_code_only = [False]
# 1-bit / ternary / 2-bit operands as fp4 (e2m1) codes on tcgen05 kind::mxf4 with unit block scales: exact, and twice the
# MAC rate of kind::i8.  QTB200_FP4=0 (or set_fp4(False)) keeps every integer operand in 8-bit lanes.
_fp4 = [os.environ.get("QTB200_FP4", "1") != "0"]


def set_fp4(flag):
    _fp4[0] = bool(flag)


def want_sign_bits(x):
    """Packed sign bits are the operand of the CUDA-core XNOR+popcount kernels only, which serve GEMV-like batches
    (rows <= POPCOUNT_MAX_M) or a forced popcount backend: larger batches skip the bit plane (and its HBM write)."""
    return x.dim() == 2 and (_force_popcount[0] or x.shape[0] <= POPCOUNT_MAX_M)


def int_codes_kind(x, bit_width=1):
    """Operand format an activation quantizer should emit for integer codes of `bit_width` bits: fp4 (e2m1) for 2-D
    inputs whose codes fit {-4..4} (sign, ternary, DoReFa-2), 8-bit lanes otherwise (conv inputs use channels-last int8)."""
    if _fp4[0] and x.dim() == 2 and bit_width <= 2:
        return L.CODES_F4
    return L.CODES_I8


class code_only_activations:
    """Context manager for fused inference chains: while active (and autograd is off) the activation quantizers
    (BinaryConnect, TernaryConnect, nnDorefaQuant / DorefaQuant, nnQuantXnor / QuantXnor) skip the fp32 fake-quant
    tensor and return a storage-less (device='meta') placeholder that only carries the low-bit operand for the next
    quantized layer.  The layer outputs are bit-identical to the default mode; what is saved is one fp32 write of
    every activation tensor.  Anything other than a quantized layer that touches the placeholder fails loudly."""

    def __enter__(self):
        self.prev = _code_only[0]
        _code_only[0] = True
        return self

    def __exit__(self, *exc):
        _code_only[0] = self.prev
        return False


_apply_grad_mode = threading.local()


def want_fp32_result(input):
    """False when the activation quantizer may skip the fp32 fake-quant tensor (code-only mode, autograd off).
    Inside autograd.Function.forward grad mode is always off, so the mode seen by TaggingFunction.apply is used."""
    grad_on = getattr(_apply_grad_mode, "value", None)
    if grad_on is None:
        grad_on = torch.is_grad_enabled()
    return not (_code_only[0] and not grad_on and input.dim() in (2, 4))


def placeholder_like(input):
    return torch.empty(input.shape, dtype=torch.float32, device="meta")


def tagged_input_device(x):
    """Device of a layer input; code-only placeholders live on 'meta' and point at their operand's device."""
    if x.is_meta:
        tag = get_tag(x)
        if tag is None:
            raise RuntimeError("pytorch_quantize_impls_b200: a code-only activation placeholder lost its low-bit operand "
                               "(it may only be passed to the next quantized layer)")
        return (tag.codes if tag.codes is not None else tag.bits).device
    ops.require_cuda(x, "input")
    return x.device


_first_layer_planes = [int(os.environ.get("QTB200_FIRST_LAYER_PLANES", "2"))]


def set_first_layer_planes(n):
    """bf16 planes of a conv layer's real-valued (fp32 image) input: 2 (default, 16 significant bits) or 3 (24 bits)."""
    if n not in (2, 3):
        raise ValueError("first-layer planes must be 2 or 3")
    _first_layer_planes[0] = n


def set_implicit_conv(flag):
    """True (default): conv layers on channels-last codes use the TMA-im2col implicit GEMM; False: explicit gather."""
    _implicit_conv[0] = bool(flag)


def set_backend(i8=None, bf16=None, popcount=None):
    """Testing/benchmark hook: force 'tcgen05' / 'simt' / 'auto' for the integer and bf16 contractions, and
    popcount=True to route every 1-bit x {1-bit, ternary} product to the XNOR-popcount kernels."""
    names = {"auto": L.BACKEND_AUTO, "tcgen05": L.BACKEND_TCGEN05, "simt": L.BACKEND_SIMT, None: None}
    if i8 is not None:
        _force_backend["i8"] = names[i8]
    if bf16 is not None:
        _force_backend["bf16"] = names[bf16]
    if popcount is not None:
        _force_popcount[0] = bool(popcount)


def get_tag(x):
    """Return the ActCodes attached to `x` by an activation quantizer if it still describes x."""
    tag = getattr(x, "_qt_codes", None)
    if tag is None:
        return None
    if tag.version != x._version or tuple(tag.shape) != tuple(x.shape):
        return None
    return tag


def attach_tag(y, tag):
    if tag is not None:
        tag.version = y._version
        tag.shape = tuple(y.shape)
        y._qt_codes = tag
    return y


class _A:
    """Activation operand handed to the contraction."""
    __slots__ = ("form", "t", "ld", "signed", "scale", "row_sum", "row_scale", "planes", "bits", "ld_bits",
                 "row_parts", "row_mul", "ready")


def _a_from_tag(tag):
    a = _A()
    a.scale, a.row_sum, a.row_scale, a.bits, a.ld_bits = tag.scale, tag.row_sum, tag.row_scale, tag.bits, tag.ld_bits
    a.row_parts, a.row_mul = tag.row_parts, tag.row_mul
    a.t, a.ld = tag.codes, tag.ld
    if tag.codes_kind in (L.CODES_I8, L.CODES_U8):
        a.form, a.signed, a.planes = "i8", tag.codes_kind == L.CODES_I8, 1
    elif tag.codes_kind == L.CODES_F4:
        a.form, a.signed, a.planes = "f4", True, 1
    elif tag.codes_kind == L.CODES_F16:
        a.form, a.signed, a.planes = "fp16", True, 1
    else:
        a.form, a.signed = "bf16", True
        a.planes = {L.CODES_BF16X2: 2, L.CODES_BF16X3: 3}.get(tag.codes_kind, 1)
    return a


def _checked_tag(tag, x):
    """The operand if its codes are inside their lane (ops.set_strict); None -> the caller contracts with the fp32 tensor on the
    real-activation route, which is what the reference computes for out-of-range DoReFa inputs."""
    if tag is None or ops.codes_in_range(tag):
        return tag
    if x.is_meta:
        raise RuntimeError("pytorch_quantize_impls_b200: a k-bit activation code left its lane (inputs of a DoReFa quantizer "
                           "outside [0, 1]) and code-only mode keeps no fp32 tensor to fall back on; clamp the activations "
                           "(Hardtanh(0, 1)) or leave code_only_activations()")
    return None


def _f4_weight_ok(pack):
    return (_force_backend["i8"] != L.BACKEND_SIMT
            and (pack.kind in ("sign", "ternary") or (pack.kind == "dorefa" and pack.bit_width <= 2)))


def _requant_i8(x2d, tag):
    """int8 codes of an already quantized fp32 tensor (quantizing twice is idempotent for sign / ternary / DoReFa)."""
    mode = {"sign": L.Q_SIGN, "ternary": L.Q_TERNARY, "dorefa": L.Q_DOREFA}[tag.kind]
    # DoReFa: y = fl(1/n) c  ->  rint(n y) = c
    _, t2 = ops.quant_act(x2d, mode, bit_width=tag.bit_width if tag.kind == "dorefa" else 0, want_y=False,
                          codes_kind=L.CODES_I8, want_bits=(tag.kind == "sign"), want_row_sum=(tag.kind == "dorefa"),
                          kind=tag.kind)
    t2.scale = tag.scale
    return _a_from_tag(t2)


def _a_split(x2d):
    """fp32 activations -> bf16 hi/mid/lo planes (3 x 8 = 24 significant bits: as faithful as the fp32 source)."""
    _, tag = ops.quant_act(x2d, L.Q_SPLIT, want_y=False, codes_kind=L.CODES_BF16X3, kind="real")
    return _a_from_tag(tag)


def _scaled(a, f):
    b = _A()
    for fld in _A.__slots__:
        if hasattr(a, fld):
            setattr(b, fld, getattr(a, fld))
    b.scale = a.scale * f
    return b


class RequantUnsupported(RuntimeError):
    """The fused requant epilogue cannot serve this call (shape / backend); callers run the unfused composition."""


class RequantSpec:
    """What a fused `layer -> [BatchNorm] -> [clamp] -> activation quantizer` chain asks of the layer's epilogue.

    mode, bit_width, kind   the quantizer (L.Q_*, DoReFa k, ActCodes.kind)
    lo, hi                  clamp in front of it (None: none)
    col_mul, col_add        folded BatchNorm affine per output column: y' = y * col_mul + col_add (None: identity)
    """

    def __init__(self, mode, kind, bit_width=0, lo=None, hi=None, col_mul=None, col_add=None):
        self.mode, self.kind, self.bit_width, self.lo, self.hi = mode, kind, bit_width, lo, hi
        self.col_mul, self.col_add = col_mul, col_add
        self.force_8bit = False      # the consumer of the codes cannot read e2m1 operands
        self.pad_channels = 0        # conv outputs: round the channel pitch of the codes up to this (0: dense) -- see conv_pad_channels
        self._fold = {}

    def fold(self, col_scale, bias, n0, n):
        """(col_scale', bias') with the BatchNorm affine folded in: y*m + a = acc*(cs*m) + (b*m + a).  Cached per operand."""
        if self.col_mul is None:
            return col_scale, bias
        key = tuple(None if t is None else (t.data_ptr(), t._version) for t in (col_scale, bias, self.col_mul, self.col_add)) + (n0, n)
        hit = self._fold.get(key)
        if hit is None:
            m, ad = self.col_mul[n0:n0 + n], self.col_add[n0:n0 + n]
            cs = m.contiguous() if col_scale is None else (col_scale * m).contiguous()
            b = ad.contiguous() if bias is None else (bias * m + ad).contiguous()
            if len(self._fold) > 8:
                self._fold.clear()
            hit = self._fold[key] = (cs, b, col_scale, bias)     # keep the sources alive: the key holds their pointers
        return hit[0], hit[1]

    def codes_kind(self, two_d):
        if self.mode == L.Q_XNOR_ROW:
            return xnor_codes_kind()
        f4 = _fp4[0] and two_d and not self.force_8bit and _force_backend["i8"] != L.BACKEND_SIMT
        if self.mode == L.Q_DOREFA:
            if f4 and self.bit_width <= 2:
                return L.CODES_F4
            return L.CODES_U8 if self.bit_width == 8 else L.CODES_I8
        return L.CODES_F4 if f4 else L.CODES_I8


def _requant_tag(spec, rq, shape, layout="rows"):
    """ActCodes describing what a requant epilogue wrote."""
    tag = ops.ActCodes()
    tag.kind, tag.bit_width = spec.kind, spec.bit_width
    tag.codes, tag.codes_kind, tag.rows, tag.cols, tag.ld = rq.codes, rq.codes_kind, rq.rows, rq.cols, rq.ld
    tag.range_ok = rq.overflow is None or ops.clamp_guarantees_lane(rq.codes_kind, spec.bit_width, spec.lo, spec.hi)
    tag.scale = 1.0
    if spec.mode == L.Q_DOREFA:
        import numpy as np
        tag.scale = float(np.float32(1.0) / np.float32(2 ** spec.bit_width - 1))
    tag.row_sum, tag.row_scale, tag.bits, tag.ld_bits = rq.row_sum_part, rq.row_part, None, 0
    tag.row_parts = rq.row_parts if (rq.row_sum_part is not None or rq.row_part is not None) else 0
    tag.row_mul = 1.0 / max(rq.cols, 1)
    tag.overflow, tag.shape, tag.version, tag.layout = rq.overflow, tuple(shape), None, layout
    return tag


def _contract(a, pack, M, N, K, out, *, w_row0=0, bias=None, out_mode=0, ldo=None, nchw_inner=1, out_offset=0,
              acc_out=None, requant=None, rq_spec=None):
    """One GEMM launch (plus the transient weight expansion) for N output columns starting at weight row w_row0.
    requant: ops.RequantOut receiving the next layer's operand (rq_spec: its RequantSpec, for the BatchNorm fold)."""
    try:
        return _contract_impl(a, pack, M, N, K, out, w_row0=w_row0, bias=bias, out_mode=out_mode, ldo=ldo,
                              nchw_inner=nchw_inner, out_offset=out_offset, acc_out=acc_out, requant=requant, rq_spec=rq_spec)
    except L.QtError as err:
        if "(code -3)" not in str(err):
            raise
        if requant is not None:
            raise RequantUnsupported(str(err))
        if a.row_parts == 0:
            raise
    # the operand came from a requant epilogue (partial row sums) but this contraction runs on a CUDA-core route (tiny or
    # unaligned shape): collapse the partial sums into plain row vectors and run again
    b = _A()
    for f in _A.__slots__:
        if hasattr(a, f):
            setattr(b, f, getattr(a, f))
    if a.row_scale is not None:
        b.row_scale = (a.row_scale[:a.row_parts].sum(0) * a.row_mul).contiguous()
    if a.row_sum is not None:
        b.row_sum = a.row_sum[:a.row_parts].sum(0).to(torch.int32).contiguous()
    b.row_parts, b.row_mul = 0, 1.0
    return _contract_impl(b, pack, M, N, K, out, w_row0=w_row0, bias=bias, out_mode=out_mode, ldo=ldo,
                          nchw_inner=nchw_inner, out_offset=out_offset, acc_out=acc_out, rq_spec=rq_spec)


def _contract_impl(a, pack, M, N, K, out, *, w_row0=0, bias=None, out_mode=0, ldo=None, nchw_inner=1, out_offset=0,
                   acc_out=None, requant=None, rq_spec=None):
    ldo = N if ldo is None else ldo
    col_scale = None if pack.col_scale is None else pack.col_scale[w_row0:w_row0 + N]
    int_w = pack.kind in ("sign", "ternary", "dorefa", "lin")
    if pack.wscale != 1.0:           # LogLin 'lin' weights: value = code * step
        a = _scaled(a, pack.wscale)
    rp = dict(row_parts=a.row_parts, row_mul=a.row_mul, requant=requant)
    if getattr(a, "ready", None) is not None:
        rp["a_ready"] = a.ready          # the operand is still being written by a quantizer on another stream (linear_overlapped)
    if requant is None and rq_spec is not None and rq_spec.lo is not None:
        rp["out_clamp"] = (rq_spec.lo, rq_spec.hi)       # a clamp activation folded behind the (BatchNorm-folded) layer

    def folded(cs, b):
        return rq_spec.fold(cs, b, w_row0, N) if rq_spec is not None else (cs, b)

    if a.form in ("i8", "f4") and int_w:
        use_pop = (a.bits is not None and pack.kind in ("sign", "ternary") and a.scale == 1.0 and out_mode == 0
                   and requant is None and rq_spec is None and a.row_parts == 0
                   and (_force_popcount[0] or (M <= POPCOUNT_MAX_M and _force_backend["i8"] == L.BACKEND_AUTO)))
        if use_pop:
            epi = ops.make_epi(out, ldo=ldo, bias=bias, scale=1.0, acc_out=acc_out, out_offset=out_offset)
            ldw = pack.ld_packed // 4
            wbits = pack.packed.view(torch.int32)
            if pack.kind == "sign":
                ops.gemm_b1b1(a.bits, a.ld_bits, wbits[0, w_row0:], ldw, M, N, K, epi)
            else:
                ops.gemm_b1t2(a.bits, a.ld_bits, wbits[0, w_row0:], wbits[1, w_row0:], ldw, M, N, K, epi)
            return
        if a.form == "f4":
            # e2m1 codes x e2m1 centred weight codes: tcgen05 kind::mxf4, unit scales, exact integer accumulators
            w, ldw = ops.expand_weight(pack, L.CODES_F4)
            cs, b = folded(col_scale, bias)
            epi = ops.make_epi(out, ldo=ldo, out_mode=out_mode, nchw_inner=nchw_inner, bias=b, col_scale=cs,
                               scale=a.scale, acc_out=acc_out, out_offset=out_offset, **rp)
            ops.gemm_f4(a.t, a.ld, w[w_row0:], ldw, M, N, K, epi)
            return
        if pack.kind == "dorefa" and pack.bit_width == 8:
            # raw unsigned codes c, W_q = (2c - 255)/255:  sum a (2c - 255) = 2 sum a c - 255 sum a
            w, ldw = ops.expand_weight(pack, L.CODES_U8)
            if a.row_sum is None:
                raise RuntimeError("internal: 8-bit DoReFa weights need activation row sums")
            cs, b = folded(col_scale, bias)
            epi = ops.make_epi(out, ldo=ldo, out_mode=out_mode, nchw_inner=nchw_inner, bias=b, col_scale=cs,
                               row_sum=a.row_sum, scale=a.scale, acc_mul=2, rs_mul=-255, acc_out=acc_out,
                               out_offset=out_offset, **rp)
            ops.gemm_i8(a.t, a.signed, a.ld, w[w_row0:], False, ldw, M, N, K, epi, _force_backend["i8"])
            return
        w, ldw = ops.expand_weight(pack, L.CODES_I8)
        cs, b = folded(col_scale, bias)
        epi = ops.make_epi(out, ldo=ldo, out_mode=out_mode, nchw_inner=nchw_inner, bias=b, col_scale=cs,
                           scale=a.scale, acc_out=acc_out, out_offset=out_offset, **rp)
        ops.gemm_i8(a.t, a.signed, a.ld, w[w_row0:], True, ldw, M, N, K, epi, _force_backend["i8"])
        return

    if a.form == "fp16":
        # XnorNet activations (+-1 / 0 codes in fp16, row_scale = mean): one fp16 tensor pass
        if int_w and pack.kind != "lin":
            w, ldw = ops.expand_weight(pack, L.CODES_F16_EXACT)
        elif pack.kind == "xnor":
            w, ldw = ops.expand_weight(pack, L.CODES_F16)
            col_scale = pack.alpha_max[w_row0:w_row0 + N]
        else:
            raise RuntimeError("internal: fp16 activation codes cannot meet a real-valued weight operand")
        cs, b = folded(col_scale, bias)
        epi = ops.make_epi(out, ldo=ldo, out_mode=out_mode, nchw_inner=nchw_inner, bias=b, col_scale=cs,
                           row_scale=a.row_scale, scale=a.scale, out_offset=out_offset, **rp)
        ops.gemm_f16(a.t, a.ld, 0, w[0, w_row0:], ldw, w.stride(0), [(0, 0)], M, N, K, epi, _force_backend["bf16"],
                     fmt=L.FMT_FP16)
        return
    if a.form != "bf16":
        raise RuntimeError("internal: integer activation codes cannot meet a real-valued weight operand")
    # bf16 routes
    if int_w or pack.kind == "log":
        w, ldw = ops.expand_weight(pack, L.CODES_BF16)       # exact in one plane
        wplanes = 1
    elif pack.kind == "xnor":
        w, ldw = ops.expand_weight(pack, L.CODES_BF16X2)
        wplanes = 2
    else:
        w, ldw, wplanes = pack.planes, pack.ld_planes, 2
    # every (A plane, W plane) product whose magnitude is above ~2^-24 of the leading term
    passes = [(i, 0) for i in range(a.planes)]
    if wplanes == 2:
        passes += [(i, 1) for i in range(min(a.planes, 2))]
    a_stride = a.t.stride(0) if a.t.dim() == 3 else 0
    w_stride = w.stride(0)
    cs, b = folded(col_scale, bias)
    epi = ops.make_epi(out, ldo=ldo, out_mode=out_mode, nchw_inner=nchw_inner, bias=b, col_scale=cs,
                       row_scale=a.row_scale, scale=a.scale, out_offset=out_offset, **rp)
    ops.gemm_f16(a.t, a.ld, a_stride, w[0, w_row0:], ldw, w_stride, passes, M, N, K, epi, _force_backend["bf16"])


def linear(x, pack, bias, requant=None, affine=None):
    """F.linear(x, W_q, bias) with W_q given as a WeightPack.  x: [..., K] fp32 CUDA tensor.
    requant (RequantSpec): fused inference chain -- the epilogue writes the next layer's low-bit operand and the
    fp32 output is never materialised; returns a code-only placeholder carrying that operand."""
    dev = tagged_input_device(x)
    K, N = pack.k, pack.n
    if x.shape[-1] != K:
        raise RuntimeError("size mismatch: input has %d features, layer expects %d" % (x.shape[-1], K))
    lead = x.shape[:-1]
    tag = get_tag(x) if x.dim() == 2 else None
    x2d = None if x.is_meta else ops.as_f32c(x).reshape(-1, K)
    tag = _checked_tag(tag, x)
    M = x.numel() // K
    if requant is not None and (x.dim() != 2 or M == 0):
        raise RequantUnsupported("fused requant needs a non-empty 2-D input")
    out = None if requant is not None else torch.empty((M, N), dtype=torch.float32, device=dev)
    if M == 0:
        return out.reshape(*lead, N)
    int_w = pack.kind in ("sign", "ternary", "dorefa", "lin")
    a = None
    if tag is not None:
        a = _a_from_tag(tag)
        if (a.form in ("i8", "f4") and not int_w) or (a.form == "fp16" and pack.kind in ("real", "lin", "log")):
            a = None
        elif a.form == "f4" and not _f4_weight_ok(pack):
            # e2m1 activation codes met weights that need 8-bit lanes (DoReFa k >= 3, or the CUDA-core backend was forced):
            # re-emit the codes as int8 from the fp32 fake-quant tensor (not available in code-only mode)
            a = None if x2d is None else _requant_i8(x2d, tag)
    if a is None:
        if x2d is None:
            raise RuntimeError("pytorch_quantize_impls_b200: this layer cannot consume the code-only activation it was "
                               "given (operand kind %r vs weight kind %r); leave code_only_activations() for this pair"
                               % (getattr(tag, "kind", None), pack.kind))
        a = _a_split(x2d)
    if bias is not None:
        bias = ops.as_f32c(bias)
    if requant is not None:
        rq = ops.RequantOut(requant.mode, requant.bit_width, requant.codes_kind(True), M, N, dev, lo=requant.lo,
                            hi=requant.hi, want_row_sum=(requant.mode == L.Q_DOREFA))
        _contract(a, pack, M, N, K, None, bias=bias, requant=rq, rq_spec=requant)
        y = torch.empty((M, N), dtype=torch.float32, device="meta")
        return attach_tag(y, _requant_tag(requant, rq, (M, N)))
    _contract(a, pack, M, N, K, out, bias=bias, rq_spec=affine)     # affine: folded BatchNorm (RequantSpec.fold), fp32 output
    return out.reshape(*lead, N)


_band_streams = {}
_banded_head = [os.environ.get("QTB200_BANDED_HEAD", "0") == "1"]


def set_banded_head(flag):
    """True: fusion.FusedActLayer runs `quantizer -> Linear` head pairs of >= 4096 rows as the two-stream band pipeline
    (linear_banded).  Default False: measured on B200 at the north-star shape the pipeline is slower than the plain pair
    (134 us vs 94 us: every band pays the prologue / tail of a persistent contraction launch, and the one quantizer CTA per SM
    that fits beside a 168-register contraction CTA reaches ~1.6 TB/s); the mechanism is kept and tested (bit-exact)."""
    _banded_head[0] = bool(flag)



def _tag_tensors(tag):
    return [t for t in (tag.codes, tag.row_sum, tag.row_scale, tag.bits, tag.overflow) if t is not None]


def linear_banded(x, quantize, pack, bias, nbands=4, affine=None):
    """`activation quantizer -> F.linear(., W_q, bias)` on an fp32 [M, K] input as a two-stream pipeline over row bands: the
    quantizer of band i+1 (a bounded grid: one 8-warp CTA per SM, no shared memory) runs on a side stream BESIDE the persistent
    tcgen05 contraction of band i, so the HBM-bound pass (read fp32, write codes) hides behind the tensor-bound one instead of
    preceding it.  The weights are expanded once.  quantize(x_rows, max_ctas) -> ActCodes (code-only).  fp32 result [M, N]."""
    dev = x.device
    M, K = x.shape
    N = pack.n
    out = torch.empty((M, N), dtype=torch.float32, device=dev)
    if bias is not None:
        bias = ops.as_f32c(bias)
    step = -(-M // nbands)
    step = -(-step // 256) * 256                      # whole CTA-pair tiles per band
    bounds = [(b0, min(M, b0 + step)) for b0 in range(0, M, step)]
    cur = torch.cuda.current_stream(dev)
    side = _band_streams.get(dev)
    if side is None:
        side = _band_streams[dev] = torch.cuda.Stream(device=dev)
    fork = torch.cuda.Event()
    fork.record(cur)
    side.wait_event(fork)
    tags, evs = [], []
    sms = ops.device_caps(dev.index if dev.index is not None else torch.cuda.current_device())["num_sms"]
    with torch.cuda.stream(side):
        for i, (b0, b1) in enumerate(bounds):
            # band 0 has the GPU to itself; the others share every SM with a contraction
            tag = quantize(x[b0:b1], 0 if i == 0 else sms)
            for t in _tag_tensors(tag):
                t.record_stream(cur)
            ev = torch.cuda.Event()
            ev.record(side)
            tags.append(tag)
            evs.append(ev)
    pack._hold_on = True
    try:
        for (b0, b1), tag, ev in zip(bounds, tags, evs):
            cur.wait_event(ev)
            a = _a_from_tag(tag)
            _contract(a, pack, b1 - b0, N, K, out[b0:b1], bias=bias, rq_spec=affine)
    finally:
        pack._hold_on, pack._hold = False, None
        cur.wait_stream(side)
    return out


_overlap_head = [os.environ.get("QTB200_OVERLAP_HEAD", "0") == "1"]


def set_overlap_head(flag):
    """True: a `sign / ternary quantizer -> quantized Linear` pair on a large fp32 input (fusion.FusedActLayer, code-only
    inference) runs (part of) the quantizer on a side stream BESIDE the contraction; the contraction's TMA producer follows
    per-row-block progress counters (QtActQuant.ready / QtEpilogue.a_ready).  False (default): one kernel after the other.
    Off by default because it is slower on B200 at the north-star shape (85.7 us plain; 105 - 127 us with 75 % - 0 % of the rows
    quantized up front): beside a 168-register contraction CTA one 8-warp quantizer CTA fits per SM, it pulls ~1.3 TB/s and
    slows the shared-memory-bound product it runs beside.  Bit-identical to the plain pair."""
    _overlap_head[0] = bool(flag)


OVERLAP_ROWS = 256          # rows per progress counter: one CTA-pair tile


def linear_overlapped(x, quantize, codes_kind, pack, bias, affine=None):
    """`activation quantizer -> F.linear(., W_q, bias)` on an fp32 [M, K] input with the two kernels running CONCURRENTLY: the
    quantizer (bounded grid: one 8-warp CTA per SM, no shared memory -- it fits beside the persistent tcgen05 CTA) walks the
    rows in order and bumps one counter per 256-row block; the contraction, launched right behind it on the calling stream,
    waits per tile for its block's counter.  The HBM-bound pass (read fp32, write codes) hides behind the shared-memory-bound
    product instead of preceding it.  quantize(x_rows, max_ctas, ready, ready_rows, codes_out) -> ActCodes (code-only, sign / ternary:
    no row vectors); codes_kind: its lane format.  fp32 result [M, N]."""
    dev = x.device
    M, K = x.shape
    N = pack.n
    if M % OVERLAP_ROWS or K % 1024:
        raise RuntimeError("internal: linear_overlapped needs M % 256 == 0 and K % 1024 == 0")
    out = torch.empty((M, N), dtype=torch.float32, device=dev)
    if bias is not None:
        bias = ops.as_f32c(bias)
    nblk = -(-M // OVERLAP_ROWS)
    flags = torch.zeros(nblk, dtype=torch.int32, device=dev)
    cur = torch.cuda.current_stream(dev)
    side = _band_streams.get(dev)
    if side is None:
        side = _band_streams[dev] = torch.cuda.Stream(device=dev)
    sms = ops.device_caps(dev.index if dev.index is not None else torch.cuda.current_device())["num_sms"]
    # Beside a 168-register contraction CTA only ONE 8-warp quantizer CTA fits per SM, and its loads in flight live in
    # registers: the co-resident quantizer pulls ~1.3 TB/s, slower than the contraction consumes rows.  So the first
    # OVERLAP_SPLIT of the rows are quantized at full speed in front of the contraction (same stream), the rest beside it.
    head = int(M * OVERLAP_SPLIT[0]) // OVERLAP_ROWS * OVERLAP_ROWS
    full = ops.codes_buffer(M, K, codes_kind, dev)
    tag = None
    try:
        if head > 0:
            tag = quantize(x[:head], 0, flags, OVERLAP_ROWS, full[:head])
        fork = torch.cuda.Event()
        fork.record(cur)                                  # x, the zeroed counters and the head rows are ready
        side.wait_event(fork)
        if head < M:
            with torch.cuda.stream(side):
                tag = quantize(x[head:], sms, flags[head // OVERLAP_ROWS:], OVERLAP_ROWS, full[head:])
            full.record_stream(side)
            flags.record_stream(side)
        tag.codes, tag.rows = full, M
        a = _a_from_tag(tag)
        a.ready = (flags, OVERLAP_ROWS, OVERLAP_ROWS * (K // 1024))      # chunk tasks per complete 256-row block
        _contract(a, pack, M, N, K, out, bias=bias, rq_spec=affine)
    finally:
        cur.wait_stream(side)
    return out


OVERLAP_SPLIT = [float(os.environ.get("QTB200_OVERLAP_SPLIT", "0.4"))]


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


_first_layer_implicit = [os.environ.get("QTB200_FIRST_LAYER_IMPLICIT", "1") != "0"]


def set_first_layer_implicit(flag):
    """True (default): conv layers on real-valued inputs with <= 8 channels (the image layers) run as an implicit GEMM on bf16
    plane pixels (qt_image_planes + qt_conv_bf16); False: explicit gather + split (qt_im2col)."""
    _first_layer_implicit[0] = bool(flag)


def conv_pad_channels(C):
    """Channel pitch a conv's channels-last codes should have for the NEXT conv's TMA im2col: K blocks of a full 128-byte swizzle
    row (C % 128 == 0) take half the TMA / MMA issues of 64-byte blocks and run the 128-byte-swizzle kernels.  C = 192 -> 256,
    576 -> 640; small or already aligned counts stay dense.  The pad channels hold zero codes (written by the producer's
    requant epilogue) and meet zero weights."""
    if C > 128 and C % 128 != 0:
        return (C + 127) // 128 * 128
    return C


def _relaid_weights(pack, key, build):
    """Small per-pack cache of re-laid conv operands (channel padding, W-fold phases), valid while the packed tensor is."""
    base = (pack.packed.data_ptr(), pack.packed._version)
    cache = pack._padw
    if cache is None or cache.get("base") != base:
        cache = pack._padw = {"base": base}
    hit = cache.get(key)
    if hit is None:
        hit = cache[key] = build()
    return hit


def _channel_padded_weights(pack, out_kind, N, taps, Cin, Cp):
    """Expanded conv operand [N, taps * Cin] re-laid as [N, taps * Cp] with zero columns for the pad channels."""
    def build():
        w, ldw = ops._expand_weight(pack, out_kind)
        wz = torch.zeros((N, taps, Cp), dtype=w.dtype, device=w.device)
        wz[:, :, :Cin] = w[:N, :taps * Cin].reshape(N, taps, Cin)
        return wz.reshape(N, taps * Cp), taps * Cp
    return _relaid_weights(pack, ("pad", out_kind, Cp), build)


_wfold = [os.environ.get("QTB200_WFOLD", "1") != "0"]


def set_wfold(flag):
    """True (default): stride-1 convs on channels-last codes with a 64-byte channel pitch read two horizontally adjacent pixels
    as one 128-byte pixel (two launches, one per output-column parity): a third fewer operand bytes pulled from L2 per output,
    which is what bounds these kernels (ResNet-18 3.149 -> 3.102 ms)."""
    _wfold[0] = bool(flag)


def _conv_wfold2(tag, pack, bias, geom, dims, requant, affine, out, residual, keep_out, need_rs, dev):
    """Implicit-GEMM conv with the W axis folded by two.  Pixels (2s, 2s+1) of a row form super pixel s of 2*Cp channels (the
    same memory).  Output column ow = 2t + p reads pixels ow - pw .. ow - pw + kw - 1, i.e. super pixels starting at
    t + floor((p - pw) / 2) with the filter shifted by r = (p - pw) mod 2 positions inside the first one: per parity p a stride-1
    conv over the super-pixel grid with its own zero-padded filter [kh, kw'_p * 2, Cp], asymmetric padding (explicit TMA corners),
    M / 2 rows, and outputs interleaved back by addressing (row pitch 2 * channels, offset p * channels)."""
    kh, kw, sh, sw, ph, pw, dh, dw, groups, OH, OW = geom
    B, Cin, Cp, H, W, O = dims
    a_signed = tag.codes_kind == L.CODES_I8
    kind = L.CODES_U8 if need_rs else L.CODES_I8
    x2 = tag.codes.view(B, H, W // 2, 2 * Cp)
    P = OH * OW
    rq0 = codes = None
    Op = O
    if requant is not None:
        ck = requant.codes_kind(False)
        Op = max(O, requant.pad_channels)
        codes = torch.empty((B, OH, OW, Op), device=dev, dtype=torch.uint8 if ck == L.CODES_U8 else torch.int8)
    # row sums of both output-column parities, parity-major (one strided copy instead of one per launch)
    rs_par = ops.patch_rowsum(tag.codes, not a_signed, geom, 0).view(-1, 2).t().contiguous() if need_rs else None
    res_flat = None if residual is None else residual.permute(0, 2, 3, 1).reshape(-1)
    cs, bg = pack.col_scale, bias
    if requant is not None or affine is not None:
        cs, bg = (requant or affine).fold(cs, bg, 0, O)
    for p in (0, 1):
        d = p - pw
        lo_w = d // 2                        # floor
        r = d - 2 * lo_w
        kwf = (r + kw - 1) // 2 + 1

        def build(r=r, kwf=kwf):
            w, _ = ops._expand_weight(pack, kind)
            wz = torch.zeros((O, kh, kwf * 2, Cp), dtype=w.dtype, device=w.device)
            wz[:, :, r:r + kw, :Cin] = w[:O, :kh * kw * Cin].reshape(O, kh, kw, Cin)
            return wz.reshape(O, kh * kwf * 2 * Cp), kh * kwf * 2 * Cp
        w, ldw = _relaid_weights(pack, ("wfold", kind, Cp, r, kw), build)
        rq = None
        if requant is not None:
            rq = ops.RequantOut(requant.mode, requant.bit_width, ck, B * P // 2, O, dev, lo=requant.lo, hi=requant.hi,
                                ld=2 * Op, codes=codes, col_offset=p * Op, cover=Op)
            if rq0 is None:
                rq0 = rq
            elif rq0.overflow is not None:
                rq.c.overflow = rq0.c.overflow           # one sticky flag for both parities
        rs = rs_par[p] if need_rs else None
        epi = ops.make_epi(out, bias=bg, col_scale=cs, row_sum=rs, acc_mul=2 if need_rs else 1, rs_mul=-255 if need_rs else 0,
                           scale=tag.scale * pack.wscale, out_offset=p * O, requant=rq, out_clamp=_out_clamp(affine, requant, keep_out),
                           out_mode=0, ldo=2 * O, nchw_inner=1,
                           residual=None if res_flat is None else res_flat[p * O:], ld_res=2 * O)
        g2 = (kh, kwf, sh, 1, ph, 0, dh, 1, 1, OH, OW // 2)
        if not ops.conv_i8(x2, a_signed, g2, 0, w, not need_rs, ldw, O, epi, corners=(lo_w, lo_w + OW // 2 - W // 2)):
            raise RuntimeError("internal: W-folded conv rejected by the implicit-GEMM kernel")
    return rq0, codes


def _out_clamp(affine, requant, keep_out):
    """Clamp applied to the fp32 value the epilogue WRITES (the requant path clamps its own copy)."""
    if affine is not None and affine.lo is not None:
        return affine.lo, affine.hi
    if keep_out and requant is not None and requant.lo is not None:
        return requant.lo, requant.hi
    return None


def _first_layer_weights(pack, O, kh, kw, Cin, P, fh, fw, khf, kwf):
    """bf16 weights of the plane-pixel implicit GEMM: [O, khf, kwf, fh, fw, 16] -- the exact integer codes of W_q repeated for
    each of the P parts (hi / mid / lo) of the input channels, zero in the unused slots and in the taps the folds add.
    Cached on the pack while its packed tensor is unchanged."""
    key = (pack.packed.data_ptr(), pack.packed._version, kh, kw, Cin, P, fh, fw)
    hit = getattr(pack, "_first", None)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    w, ldw = ops._expand_weight(pack, L.CODES_BF16)             # [1, O, ld]: exact codes, K order (kh, kw, c)
    wv = w[0, :O, :kh * kw * Cin].reshape(O, kh, kw, Cin)
    wz = torch.zeros((O, khf * fh, kwf * fw, 16), dtype=torch.bfloat16, device=w.device)
    for p in range(P):
        wz[:, :kh, :kw, p * Cin:(p + 1) * Cin] = wv
    # (ky' fh + dy, kx' fw + dx) -> [ky', kx', dy, dx]: the K order of a super pixel
    wz = wz.reshape(O, khf, fh, kwf, fw, 16).permute(0, 1, 3, 2, 4, 5).reshape(O, khf * kwf * fh * fw * 16).contiguous()
    pack._first = (key, wz, wz.shape[1])
    return wz, wz.shape[1]


_first_layer_windows = [os.environ.get("QTB200_FIRST_LAYER_WINDOWS", "1") != "0"]


def set_first_layer_windows(flag):
    """True (default): image layers whose filter row fits one record (kw * parts * Cin <= 128 slots) read row-window records
    (qt_image_windows) with one or two k-blocks per filter row; False: plane pixels with the space-to-depth folds only."""
    _first_layer_windows[0] = bool(flag)


def _window_weights(pack, O, kh, kw, Cin, P, slots):
    """bf16 weights against row-window records: [O, kh, slots], slot kx * P * Cin + p * Cin + c = code of W_q[o, c, ky, kx] for
    every part p, zero in the unused slots.  Cached on the pack while its packed tensor is unchanged."""
    key = (pack.packed.data_ptr(), pack.packed._version, kh, kw, Cin, P, "win", slots)
    hit = getattr(pack, "_first", None)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    w, ldw = ops._expand_weight(pack, L.CODES_BF16)             # [1, O, ld]: exact codes, K order (kh, kw, c)
    wv = w[0, :O, :kh * kw * Cin].reshape(O, kh, kw, 1, Cin)
    wz = torch.zeros((O, kh, slots), dtype=torch.bfloat16, device=w.device)
    wz[:, :, :kw * P * Cin] = wv.expand(O, kh, kw, P, Cin).reshape(O, kh, kw * P * Cin)
    wz = wz.reshape(O, kh * slots).contiguous()
    pack._first = (key, wz, wz.shape[1])
    return wz, wz.shape[1]


def _conv_first_layer(xf, pack, geom, O, Cin, epi_kw):
    """Real-valued NCHW input with few channels: one pass over the image builds zero-padded channels-last bf16 plane pixels
    (each channel as hi / mid / lo bf16 parts side by side, 32 bytes per pixel), and the conv runs as an implicit GEMM fed by TMA
    im2col.  Strides are folded into the pixel (space-to-depth): stride 2 on both axes -> 2 x 2 super pixels of 128 bytes and a
    stride-1 filter of floor((k - 1) / 2) + 1 taps per axis; stride 4 -> 1 x 4 super pixels.  Returns False when the shape does
    not fit (the caller gathers explicitly)."""
    kh, kw, sh, sw, ph, pw, dh, dw, groups, OH, OW = geom
    if groups != 1 or Cin > 8 or pack.kind not in ("sign", "ternary", "dorefa", "lin") or pack.packed is None:
        return False
    if pack.kind == "dorefa" and not (1 <= pack.bit_width <= 8):
        return False
    if sh > 8 or sw > 8:
        return False
    B = xf.shape[0]
    P = 3 if 3 * Cin <= 16 else 2
    if _first_layer_windows[0] and dw == 1 and kw * P * Cin <= 128 and kw > 1:
        # row-window records: K = kh records (bytes per output pixel: kh * slots * 2; 128 slots = two 128-byte k-blocks per
        # filter row: AlexNet's 11x11 / 4 reads 11 x 256 B per output instead of 33 x 128 B in the 1 x 4 space-to-depth form)
        used = kw * P * Cin
        slots = 16 if used <= 16 else (32 if used <= 32 else (64 if used <= 64 else 128))
        Hp = (OH - 1) * sh + (kh - 1) * dh + 1
        rec = ops.image_windows(xf, P, kw, sw, ph, pw, Hp, OW, slots)
        wz, ldw = _window_weights(pack, O, kh, kw, Cin, P, slots)
        epi = ops.make_epi(**epi_kw)
        ops.conv_bf16(rec, (B, slots, Hp, OW, kh, 1, sh, 1, 0, 0, dh, 1, 1, 0, OH, OW), wz, ldw, O, epi)
        return True
    fw = sw if (sw in (2, 4) and dw == 1) else 1
    fh = 2 if (sh == 2 and dh == 1 and fw == 2) else 1          # 2 x 2 x 32 B = one 128-byte swizzle row
    if fw > 1:
        kwf = (kw - 1) // fw + 1
        Wp = fw * (OW + kwf - 1)
        g_w, g_kw, g_sw, g_dw = Wp // fw, kwf, 1, 1
    else:
        kwf = kw
        Wp = (OW - 1) * sw + (kw - 1) * dw + 1
        g_w, g_kw, g_sw, g_dw = Wp, kw, sw, dw
    if fh > 1:
        khf = (kh - 1) // fh + 1
        Hp = fh * (OH + khf - 1)
        g_h, g_kh, g_sh, g_dh = Hp // fh, khf, 1, 1
    else:
        khf = kh
        Hp = (OH - 1) * sh + (kh - 1) * dh + 1
        g_h, g_kh, g_sh, g_dh = Hp, kh, sh, dh
    planes = ops.image_planes(xf, P, ph, pw, Hp, Wp, fh, fw)
    wz, ldw = _first_layer_weights(pack, O, kh, kw, Cin, P, fh, fw, khf, kwf)
    epi = ops.make_epi(**epi_kw)
    ops.conv_bf16(planes, (B, 16 * fh * fw, g_h, g_w, g_kh, g_kw, g_sh, g_sw, 0, 0, g_dh, g_dw, 1, 0, OH, OW), wz, ldw, O, epi)
    return True


def conv2d(x, pack, bias, weight_shape, stride, padding, dilation, groups, requant=None, affine=None, out_format="nchw",
           residual=None, keep_out=False):
    """F.conv2d(x, W_q, bias, stride, padding, dilation, groups): implicit GEMM (TMA im2col) on channels-last codes or bf16
    plane pixels, or an explicit gather + GEMM for the shapes those decline.
    requant (RequantSpec): the epilogue writes the next conv's channels-last codes [B, OH, OW, O] instead of the fp32 tensor
    (raises RequantUnsupported when the call cannot take an implicit-GEMM route); keep_out=True writes the fp32 tensor too.
    out_format "nhwc": the fp32 result is a channels-last tensor (memory [B, OH, OW, O], full-line TMA stores);
    residual: channels-last fp32 [B, O, OH, OW] added before the clamp / requant (out_format "nhwc" only)."""
    dev = tagged_input_device(x)
    if x.dim() != 4:
        raise RuntimeError("expected a 4-D NCHW input, got %d-D" % x.dim())
    B, Cin, H, W = x.shape
    O, Cg, kh, kw = weight_shape
    if Cin != Cg * groups:
        raise RuntimeError("input has %d channels, layer expects %d" % (Cin, Cg * groups))
    sh, sw = _pair(stride)
    dh, dw = _pair(dilation)
    if isinstance(padding, str):
        if padding == "valid":
            ph = pw = 0
        else:
            tot_h, tot_w = dh * (kh - 1), dw * (kw - 1)
            if padding != "same" or tot_h % 2 or tot_w % 2 or sh != 1 or sw != 1:
                raise NotImplementedError("padding=%r is not supported by the quantized conv kernels" % (padding,))
            ph, pw = tot_h // 2, tot_w // 2
    else:
        ph, pw = _pair(padding)
    OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    if requant is not None and (B == 0 or OH <= 0 or OW <= 0 or groups != 1 or O % 32 != 0 or requant.mode == L.Q_XNOR_ROW):
        raise RequantUnsupported("fused conv requant needs groups == 1, out_channels % 32 == 0 and a non-empty output")
    nhwc_out = out_format == "nhwc"
    if residual is not None:
        if not nhwc_out or not ops.is_channels_last(residual) or tuple(residual.shape) != (B, O, OH, OW) \
                or residual.dtype != torch.float32:
            raise RuntimeError("internal: the residual must be a channels-last fp32 tensor of the output's shape")
    want_out = requant is None or keep_out
    out = None
    if want_out:
        out = torch.empty((B, O, max(OH, 0), max(OW, 0)), dtype=torch.float32, device=dev,
                          memory_format=torch.channels_last if nhwc_out else torch.contiguous_format)
    if B == 0 or OH <= 0 or OW <= 0:
        return out
    # output addressing of the epilogue: NCHW (per-column coalesced stores) or row-major [pixels, channels] (TMA stores)
    out_kw = (dict(out_mode=0, ldo=O, nchw_inner=1) if nhwc_out else dict(out_mode=1, ldo=O, nchw_inner=OH * OW))
    if residual is not None:
        out_kw.update(residual=residual, ld_res=O)
    Ng, Kg, P = O // groups, Cg * kh * kw, OH * OW
    geom = (kh, kw, sh, sw, ph, pw, dh, dw, groups, OH, OW)
    int_w = pack.kind in ("sign", "ternary", "dorefa", "lin")
    tag = get_tag(x)
    if tag is not None and not (tag.codes_kind in (L.CODES_I8, L.CODES_U8) and int_w and tag.layout == "nhwc"):
        tag = None
    tag = _checked_tag(tag, x)
    if bias is not None:
        bias = ops.as_f32c(bias)
    need_rs = int_w and pack.kind == "dorefa" and pack.bit_width == 8
    if x.is_meta and tag is None:
        raise RuntimeError("pytorch_quantize_impls_b200: this conv layer cannot consume the code-only activation it was given")

    # implicit GEMM: TMA im2col straight from the channels-last codes (no im2col matrix is materialised)
    Cp = tag.codes.shape[3] if (tag is not None and tag.codes.dim() == 4) else Cin       # channel pitch of the codes (>= Cin)
    if tag is not None and Cp != Cin and (groups != 1 or not _implicit_conv[0] or _force_backend["i8"] == L.BACKEND_SIMT):
        raise RuntimeError("internal: channel-padded codes need the implicit-GEMM route (groups == 1)")
    if (tag is not None and _implicit_conv[0] and Cg % 32 == 0 and Cin % 16 == 0
            and _force_backend["i8"] != L.BACKEND_SIMT):
        a_signed = tag.codes_kind == L.CODES_I8
        # DoReFa-8 weights stay unsigned codes c; the zero point needs the per-pixel patch sums of the activation codes
        if (_wfold[0] and groups == 1 and Cp == 64 and sw == 1 and dw == 1 and W % 2 == 0 and OW % 2 == 0 and kw <= 7
                and (out is None or nhwc_out) and B * P >= 148 * 256):
            rq, codes = _conv_wfold2(tag, pack, bias, geom, (B, Cin, Cp, H, W, O), requant, affine, out, residual, keep_out, need_rs, dev)
            if requant is not None:
                y = out if keep_out else torch.empty((B, O, OH, OW), dtype=torch.float32, device="meta")
                t = _requant_tag(requant, rq, (B, O, OH, OW), layout="nhwc")
                t.codes, t.rows, t.cols, t.ld = codes, B, codes.shape[3] * P, codes.shape[3] * P
                return attach_tag(y, t)
            return out
        if Cp != Cin:
            w, ldw = _channel_padded_weights(pack, L.CODES_U8 if need_rs else L.CODES_I8, O, kh * kw, Cin, Cp)
        else:
            w, ldw = ops.expand_weight(pack, L.CODES_U8 if need_rs else L.CODES_I8)
        col_scale = pack.col_scale
        done = True
        rq = None
        if requant is not None:
            Op = max(O, requant.pad_channels) if groups == 1 else O        # channel pitch of the codes (pad channels: zero codes)
            rq = ops.RequantOut(requant.mode, requant.bit_width, requant.codes_kind(False), B * P, O, dev, lo=requant.lo,
                                hi=requant.hi, ld=Op,
                                codes=torch.empty((B, OH, OW, Op), device=dev,
                                                  dtype=torch.uint8 if requant.codes_kind(False) == L.CODES_U8 else torch.int8))
        for g in range(groups):
            rs = ops.patch_rowsum(tag.codes, not a_signed, geom, g) if need_rs else None
            cs = None if col_scale is None else col_scale[g * Ng:(g + 1) * Ng]
            bg = None if bias is None else bias[g * Ng:(g + 1) * Ng]
            if requant is not None or affine is not None:
                cs, bg = (requant or affine).fold(cs, bg, g * Ng, Ng)
            epi = ops.make_epi(out, bias=bg, col_scale=cs,
                               row_sum=rs, acc_mul=2 if need_rs else 1, rs_mul=-255 if need_rs else 0,
                               scale=tag.scale * pack.wscale, out_offset=g * Ng * (1 if nhwc_out else P), requant=rq,
                               out_clamp=_out_clamp(affine, requant, keep_out),
                               **out_kw)
            if not ops.conv_i8(tag.codes, a_signed, geom, g, w[g * Ng:], not need_rs, ldw, Ng, epi):
                done = False
                break
        if done:
            if requant is not None:
                y = out if keep_out else torch.empty((B, O, OH, OW), dtype=torch.float32, device="meta")
                return attach_tag(y, _requant_tag(requant, rq, (B, O, OH, OW), layout="nhwc"))
            return out

    # real-valued image input (first layers): implicit GEMM on bf16 plane pixels
    if tag is None and not x.is_meta and int_w and _first_layer_implicit[0] and _force_backend["bf16"] != L.BACKEND_SIMT \
            and (requant is None or (groups == 1 and O % 32 == 0)):
        rq = None
        if requant is not None:
            ck = requant.codes_kind(False)
            Op = max(O, requant.pad_channels)
            rq = ops.RequantOut(requant.mode, requant.bit_width, ck, B * P, O, dev, lo=requant.lo, hi=requant.hi, ld=Op,
                                codes=torch.empty((B, OH, OW, Op), device=dev,
                                                  dtype=torch.uint8 if ck == L.CODES_U8 else torch.int8))
        cs, bg = pack.col_scale, bias
        if requant is not None or affine is not None:
            cs, bg = (requant or affine).fold(cs, bg, 0, O)
        epi_kw = dict(out=out, bias=bg, col_scale=cs, scale=pack.wscale, requant=rq,
                      out_clamp=_out_clamp(affine, requant, keep_out), **out_kw)
        if _conv_first_layer(ops.as_f32c(x), pack, geom, O, Cin, epi_kw):
            if requant is not None:
                y = out if keep_out else torch.empty((B, O, OH, OW), dtype=torch.float32, device="meta")
                return attach_tag(y, _requant_tag(requant, rq, (B, O, OH, OW), layout="nhwc"))
            return out
    if requant is not None:
        raise RequantUnsupported("fused conv requant needs an implicit-GEMM route (channels-last 8-bit codes in with C/groups % 32 == 0, "
                                 "or an image input with <= 8 channels)")
    if nhwc_out:
        # the explicit-gather route below writes NCHW: run it and convert (rare shapes only)
        aff = affine
        if residual is not None and affine is not None and affine.lo is not None:      # the clamp comes after the add
            aff = RequantSpec(-1, None, col_mul=affine.col_mul, col_add=affine.col_add)
        y = conv2d(x, pack, bias, weight_shape, stride, padding, dilation, groups, affine=aff)
        if residual is not None:
            y = y + residual
            if affine is not None and affine.lo is not None:
                y = torch.clamp(y, affine.lo, affine.hi)
        return y.contiguous(memory_format=torch.channels_last)

    xf = None
    if tag is not None:
        elem, ld, nplanes = 1, ops.round_up(Kg, 16), 1
        dtype = tag.codes.dtype
    else:
        # real-valued input (first layers): fused gather + bf16 split straight from the fp32 NCHW tensor.  Two planes (hi/lo,
        # 16 significant bits: relative error <= 2^-17 per element, two orders of magnitude inside the 1e-3 tolerance) by
        # default -- the im2col planes of a 224x224 first conv are GBs, a third plane costs 50 % more traffic and MMA passes;
        # set_first_layer_planes(3) restores the fp32-faithful hi/mid/lo split
        elem, ld, nplanes = 2, ops.round_up(Kg, 8), _first_layer_planes[0]
        xf = ops.as_f32c(x)
        dtype = torch.bfloat16

    # bound the transient im2col matrix (~1.5 GiB per chunk of images)
    per_img = P * ld * elem * nplanes
    bchunk = max(1, min(B, (3 << 29) // max(per_img, 1)))
    for b0 in range(0, B, bchunk):
        b1 = min(B, b0 + bchunk)
        nb = b1 - b0
        M = nb * P
        for g in range(groups):
            a = _A()
            a.bits, a.ld_bits, a.row_scale, a.row_sum = None, 0, None, None
            a.row_parts, a.row_mul = 0, 1.0
            a.ld = ld
            buf = torch.empty((nplanes, M, ld), dtype=dtype, device=dev)
            if tag is not None:
                if need_rs:
                    a.row_sum = torch.empty(M, dtype=torch.int32, device=dev)
                ops.im2col(tag.codes[b0:b1], 1, geom, g, buf[0], ld, row_sum=a.row_sum,
                           is_unsigned=(dtype == torch.uint8), nhwc=True)
                a.form, a.t, a.signed, a.scale, a.planes = "i8", buf[0], dtype == torch.int8, tag.scale, 1
            else:
                ops.im2col(xf[b0:b1], 4, geom, g, buf, ld, split3=nplanes)
                a.form, a.t, a.signed, a.scale, a.planes = "bf16", buf, True, 1.0, nplanes
            _contract(a, pack, M, Ng, Kg, out, w_row0=g * Ng, bias=None if bias is None else bias[g * Ng:(g + 1) * Ng],
                      out_mode=1, ldo=O, nchw_inner=P, out_offset=(b0 * O + g * Ng) * P, rq_spec=affine)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# backward: gradient contractions of the dense layers on the bf16 tensor-core route
# ---------------------------------------------------------------------------------------------------------------------
_grad_backend = [os.environ.get("QTB200_GRAD_GEMM", "tcgen05")]


def set_grad_backend(name):
    """'tcgen05' (default): grad_x / grad_W of the dense layers run as bf16 hi/lo plane products on qt_gemm_f16 (relative error
    <= 2^-16, inside the 1e-4 the gradient parity tests ask for); 'torch': fp32 torch.matmul."""
    if name not in ("tcgen05", "torch"):
        raise ValueError("grad backend must be 'tcgen05' or 'torch'")
    _grad_backend[0] = name


_GRAD_PASSES = [(0, 0), (1, 0), (0, 1)]        # hi*hi + lo*hi + hi*lo  (lo*lo is below 2^-32 of the leading term)


def grad_input_linear(g2d, wq2d):
    """grad_x[m, k] = sum_n g[m, n] * W_q[n, k]   (g . W_q, e.g. binary_connect.py:104-105).  fp32 in, fp32 out."""
    if _grad_backend[0] != "tcgen05" or not g2d.is_cuda:
        return g2d @ wq2d
    g2d, wq2d = ops.as_f32c(g2d), ops.as_f32c(wq2d)
    M, N = g2d.shape
    K = wq2d.shape[1]
    out = torch.empty((M, K), dtype=torch.float32, device=g2d.device)
    if M == 0 or K == 0 or N == 0:
        return out.zero_()
    _, tag = ops.quant_act(g2d, L.Q_SPLIT, want_y=False, codes_kind=L.CODES_BF16X2, kind="real")     # [2, M, ld]
    wT, ldw = ops.transpose_split(wq2d, planes=2)                                                     # [2, K, ld]
    epi = ops.make_epi(out, ldo=K)
    ops.gemm_f16(tag.codes, tag.ld, tag.codes.stride(0), wT, ldw, wT.stride(0), _GRAD_PASSES, M, K, N, epi,
                 _force_backend["bf16"])
    return out


def grad_weight_linear(g2d, x2d):
    """grad_Wq[n, k] = sum_m g[m, n] * x[m, k]   (g^T . x, e.g. binary_connect.py:106-107): the reduction runs over the batch, so
    both operands are transposed + split in one pass each."""
    if _grad_backend[0] != "tcgen05" or not g2d.is_cuda:
        return g2d.t() @ x2d
    g2d, x2d = ops.as_f32c(g2d), ops.as_f32c(x2d)
    M, N = g2d.shape
    K = x2d.shape[1]
    out = torch.empty((N, K), dtype=torch.float32, device=g2d.device)
    if M == 0 or K == 0 or N == 0:
        return out.zero_()
    gT, ldg = ops.transpose_split(g2d, planes=2)          # [2, N, ld(M)]
    xT, ldx = ops.transpose_split(x2d, planes=2)          # [2, K, ld(M)]
    epi = ops.make_epi(out, ldo=K)
    ops.gemm_f16(gT, ldg, gT.stride(0), xT, ldx, xT.stride(0), _GRAD_PASSES, N, K, M, epi, _force_backend["bf16"])
    return out


def _conv_geom(input_shape, weight_shape, stride, padding, dilation):
    B, Cin, H, W = input_shape
    O, Cg, kh, kw = weight_shape
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    return B, Cin, H, W, O, Cg, kh, kw, (sh, sw), (ph, pw), (dh, dw), OH, OW


def grad_input_conv2d(input_shape, wq, grad_output, stride=1, padding=0, dilation=1, groups=1):
    """d loss / d x of F.conv2d(x, W_q): the transposed convolution as a contraction over the output channels,
    cols[m, (c, kh, kw)] = sum_o g[m, o] W_q[o, (c, kh, kw)] on the bf16 hi/lo tensor-core route (grad_input_linear), followed by
    the col2im scatter-add (torch fold: pure data movement).  binary_connect.py:139-142, terner_connect.py:137-139,
    dorefa_connect.py:181-183, xnor_connect.py:152-154 call torch.nn.grad.conv2d_input for this."""
    if _grad_backend[0] != "tcgen05" or not grad_output.is_cuda or isinstance(padding, str):
        return torch.nn.grad.conv2d_input(input_shape, wq, grad_output, stride=stride, padding=padding, dilation=dilation, groups=groups)
    B, Cin, H, W, O, Cg, kh, kw, st, pd, dl, OH, OW = _conv_geom(input_shape, wq.shape, stride, padding, dilation)
    Ng, L_ = O // groups, OH * OW
    g = ops.as_f32c(grad_output)
    parts = []
    for gi in range(groups):
        g2 = g[:, gi * Ng:(gi + 1) * Ng].permute(0, 2, 3, 1).reshape(B * L_, Ng)                 # [M, Ng]
        w2 = ops.as_f32c(wq[gi * Ng:(gi + 1) * Ng]).reshape(Ng, Cg * kh * kw)
        cols = grad_input_linear(g2.contiguous(), w2)                                               # [M, Cg*kh*kw]
        parts.append(cols.reshape(B, L_, Cg * kh * kw).transpose(1, 2))
    cols = parts[0] if groups == 1 else torch.cat(parts, 1)
    return torch.nn.functional.fold(cols, (H, W), (kh, kw), dilation=dl, padding=pd, stride=st)


def grad_weight_conv2d(x, weight_shape, grad_output, stride=1, padding=0, dilation=1, groups=1):
    """d loss / d W_q of F.conv2d(x, W_q): gw[o, (c, kh, kw)] = sum_m g[m, o] im2col(x)[m, (c, kh, kw)] -- the reduction runs over
    the B*OH*OW output pixels, both operands are transposed + split in one pass each (grad_weight_linear).  The im2col view comes
    from torch unfold (data movement).  binary_connect.py:143-146 etc. call torch.nn.grad.conv2d_weight for this."""
    if _grad_backend[0] != "tcgen05" or not grad_output.is_cuda or isinstance(padding, str):
        return torch.nn.grad.conv2d_weight(x, weight_shape, grad_output, stride=stride, padding=padding, dilation=dilation, groups=groups)
    B, Cin, H, W, O, Cg, kh, kw, st, pd, dl, OH, OW = _conv_geom(x.shape, weight_shape, stride, padding, dilation)
    Ng, L_ = O // groups, OH * OW
    g = ops.as_f32c(grad_output)
    unf = torch.nn.functional.unfold(ops.as_f32c(x), (kh, kw), dilation=dl, padding=pd, stride=st)   # [B, Cin*kh*kw, L]
    outs = []
    for gi in range(groups):
        g2 = g[:, gi * Ng:(gi + 1) * Ng].permute(0, 2, 3, 1).reshape(B * L_, Ng).contiguous()
        cx = unf[:, gi * Cg * kh * kw:(gi + 1) * Cg * kh * kw].transpose(1, 2).reshape(B * L_, Cg * kh * kw).contiguous()
        outs.append(grad_weight_linear(g2, cx))                                                        # [Ng, Cg*kh*kw]
    gw = outs[0] if groups == 1 else torch.cat(outs, 0)
    return gw.reshape(weight_shape)
