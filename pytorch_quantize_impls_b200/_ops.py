"""Tensor-level wrappers over the C ABI: torch is used only for device memory and streams."""
import ctypes as C

import torch

from . import _lib as L

_strict = False


def set_strict(flag):
    """Policy for k-bit activation codes that leave their 8-bit (or e2m1) lane -- the reference's _quantize does not clamp
    (dorefa_connect.py:24-25), so inputs outside [0, 1] produce codes outside [0, 2^k - 1], and far enough outside they no
    longer fit the lane (int8 for k <= 7, uint8 for k = 8, {-4..4} for the e2m1 lane of k = 2):

      False (default)  the quantizer saturates the code and raises a sticky device flag; the layer that CONSUMES the operand
                       checks the flag before contracting (one host synchronisation per quantizer whose range is not
                       guaranteed by a fused clamp) and, when it is set, contracts with the fp32 fake-quant tensor on the
                       real-activation route instead -- the output then equals the reference's; in code-only mode, where
                       no fp32 tensor exists, it raises
      True             synchronise and raise right after every k-bit quantizer
      "off"            never look at the flag (CUDA-graph capture behaves like this: a capture cannot synchronise); the
                       caller promises inputs in lane range, e.g. [0, 1] behind a Hardtanh(0, 1)
    """
    global _strict
    if flag not in (True, False, "off"):
        raise ValueError("set_strict: True, False or 'off'")
    _strict = flag


def codes_in_range(tag):
    """True when the consumer may contract on `tag`'s codes (see set_strict).  Caches a clean result on the tag."""
    if tag is None or tag.overflow is None or tag.range_ok or _strict == "off":
        return True
    if torch.cuda.is_current_stream_capturing():
        return True
    ok = int(tag.overflow.item()) == 0
    if ok:
        tag.range_ok = True
    return ok


def _lane_range(codes_kind):
    return {L.CODES_I8: (-128, 127), L.CODES_U8: (0, 255), L.CODES_F4: (-4, 4)}.get(codes_kind)


def clamp_guarantees_lane(codes_kind, bit_width, lo, hi):
    """A clamp to [lo, hi] in front of a DoReFa-k quantizer keeps round(n x) inside the lane of `codes_kind`."""
    r = _lane_range(codes_kind)
    if r is None or lo is None or hi is None:
        return r is None
    n = float(2 ** bit_width - 1)
    return round(n * lo) >= r[0] and round(n * hi) <= r[1]


def _overflow_flag(dev, mode, guaranteed=False):
    """Sticky lane-overflow flag of a DoReFa quantizer, or None when nobody will look at it: other quantizers, a clamp that
    keeps the codes inside the lane, or set_strict("off") -- one memset launch per quantizer saved on the inference chains."""
    if mode != L.Q_DOREFA or guaranteed or _strict == "off":
        return None
    return torch.zeros(1, dtype=torch.int32, device=dev)


def round_up(a, b):
    return (a + b - 1) // b * b


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(
            "pytorch_quantize_impls_b200: %s must be a CUDA tensor (got %s); the quantized kernels are "
            "sm_100a CUDA code and there is no CPU fallback" % (what, getattr(t, "device", type(t))))


def as_f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class ActCodes:
    """Low-bit activation operand produced by an activation quantizer and consumed by the next layer.

    kind        'sign' | 'ternary' | 'dorefa' | 'xnor'
    codes       int8/uint8 [rows, ld], bf16/fp16 [rows, ld] or fp4 (e2m1, uint8 [rows, ld/2]) tensor (K-major, zero padded)
    scale       value = scale * code  (DoReFa: fl(1/n); others 1.0)
    row_sum     int32 [rows] sum of codes (for unsigned-weight zero points), or None
    row_scale   fp32 [rows] (XnorNet row mean), or None
    bits        uint32 [rows, ld_bits] packed signs (kind 'sign'), or None
    shape       shape of the fp32 tensor the codes describe (e.g. NCHW)
    """
    __slots__ = ("kind", "bit_width", "codes", "codes_kind", "rows", "cols", "ld", "scale", "row_sum",
                 "row_scale", "bits", "ld_bits", "overflow", "shape", "version", "layout", "row_parts", "row_mul",
                 "range_ok")
    # range_ok: the codes are known to sit inside their lane (a clamp in front of the quantizer, or a checked clean flag)
    # row_parts > 0: the operand was written by a fused requant epilogue; row_scale / row_sum are [row_parts, rows]
    # partial sums and the consumer epilogue uses  row_mul * sum_p row_scale[p]  /  sum_p row_sum[p]

    def __init__(self):
        self.range_ok = False

    def check(self):
        if self.overflow is not None and int(self.overflow.item()) != 0:
            raise RuntimeError("quantized activation code overflowed its 8-bit lane "
                               "(inputs of a k-bit DoReFa quantizer must lie in [0, 1])")


def quant_act(x, mode, *, bit_width=0, fsr=0, with_sign=1, want_y=True, codes_kind=L.CODES_NONE,
              want_bits=False, want_row_sum=False, want_row_scale=False, kind=None, pre=None, max_ctas=0, ready=None,
              ready_rows=0, codes_out=None):
    """Run one activation-quantizer pass.  Returns (y or None, ActCodes or None).
    pre = (scale[C], shift[C], lo, hi) fuses x' = clamp(x*scale[ch] + shift[ch], lo, hi) in front (lo/hi None: no clamp)."""
    require_cuda(x, "input")
    x = as_f32c(x)
    shape = tuple(x.shape)
    rows = shape[0] if x.dim() == 2 else 1      # N-D tensors (NCHW activations) are one dense row
    cols = x.numel() // max(rows, 1) if rows else 0
    dev = x.device
    y = torch.empty_like(x) if want_y else None
    a = L.QtActQuant()
    a.mode, a.bit_width, a.fsr, a.with_sign = mode, bit_width, fsr, int(with_sign)
    a.x, a.rows, a.cols, a.ld_x = _p(x), rows, cols, cols
    a.y, a.ld_y = _p(y), cols
    if pre is not None:
        ps, pt, lo, hi = pre
        a.pre_scale, a.pre_shift, a.pre_channels = _p(ps), _p(pt), ps.numel()
        a.pre_hw = 1 if x.dim() == 2 else (x.numel() // (x.shape[0] * x.shape[1]))
        a.pre_clamp = 0 if lo is None else 1
        a.pre_lo, a.pre_hi = (0.0, 0.0) if lo is None else (float(lo), float(hi))
    codes = bits = row_sum = row_scale = overflow = None
    ld = ldb = 0
    layout = "rows"
    if x.dim() == 4 and codes_kind in (L.CODES_I8, L.CODES_U8) and not want_bits and not want_row_sum:
        # conv activations: channels-last codes [B, H, W, C] (what the conv gather reads with 16-byte vectors)
        layout = "nhwc"
        Bn, Cn, Hn, Wn = shape
        rows, cols = Bn, Cn * Hn * Wn
        a.rows, a.cols, a.ld_x, a.ld_y, a.nhwc_c = rows, cols, cols, cols, Cn
        codes = torch.empty((Bn, Hn, Wn, Cn), dtype=torch.int8 if codes_kind == L.CODES_I8 else torch.uint8, device=dev)
        overflow = _overflow_flag(dev, mode)
        ld = cols
    elif codes_kind in (L.CODES_I8, L.CODES_U8):
        ld = round_up(max(cols, 1), 16)
        codes = torch.empty((rows, ld), dtype=torch.int8 if codes_kind == L.CODES_I8 else torch.uint8, device=dev)
        overflow = _overflow_flag(dev, mode)
    elif codes_kind == L.CODES_F4:
        ld = round_up(max(cols, 1), 32)        # elements; two e2m1 codes per byte -> 16-byte rows
        codes = torch.empty((rows, ld // 2), dtype=torch.uint8, device=dev)
        overflow = _overflow_flag(dev, mode)
    elif codes_kind == L.CODES_BF16:
        ld = round_up(max(cols, 1), 8)
        codes = torch.empty((rows, ld), dtype=torch.bfloat16, device=dev)
    elif codes_kind == L.CODES_F16:
        ld = round_up(max(cols, 1), 8)
        codes = torch.empty((rows, ld), dtype=torch.float16, device=dev)
    elif codes_kind in (L.CODES_BF16X2, L.CODES_BF16X3):
        ld = round_up(max(cols, 1), 8)
        codes = torch.empty((2 if codes_kind == L.CODES_BF16X2 else 3, rows, ld), dtype=torch.bfloat16, device=dev)
    if codes_out is not None:            # row range of a larger operand (codes_buffer): 2-D code matrices only
        if layout != "rows" or codes is None or codes_out.shape != codes.shape or codes_out.dtype != codes.dtype:
            raise ValueError("quant_act: codes_out does not match the operand this call produces")
        codes = codes_out
    if want_bits:
        ldb = round_up((cols + 31) // 32, 4)
        bits = torch.empty((rows, ldb), dtype=torch.int32, device=dev)
    if want_row_sum:
        row_sum = torch.empty(rows, dtype=torch.int32, device=dev)
    row_parts = 0
    if want_row_scale:
        if mode == L.Q_XNOR_ROW and not want_y and x.dim() == 2:
            # one-pass form: long rows are cut into 1024-column chunks that each leave a partial row sum
            row_parts = int(L.lib().qt_quant_xnor_parts(cols, 0, (cols + 1023) // 1024))
        if row_parts > 1:
            row_scale = torch.empty((row_parts, rows), dtype=torch.float32, device=dev)
            a.row_parts = row_parts
        else:
            row_parts = 0
            row_scale = torch.empty(rows, dtype=torch.float32, device=dev)
    a.max_ctas = int(max_ctas)
    if ready is not None:                # progress counters for a consumer running beside this call (QtActQuant.ready)
        a.ready, a.ready_rows = _p(ready), int(ready_rows)
    a.codes, a.codes_kind, a.ld_codes = _p(codes), codes_kind, ld
    a.bits, a.ld_bits = _p(bits), ldb
    a.row_sum, a.row_scale, a.overflow = _p(row_sum), _p(row_scale), _p(overflow)
    if rows and cols:
        L.check(L.lib().qt_quant_act(C.byref(a), _stream()), "qt_quant_act")
    tag = None
    if codes is not None or bits is not None:
        tag = ActCodes()
        tag.kind, tag.bit_width = kind, bit_width
        tag.codes, tag.codes_kind, tag.rows, tag.cols, tag.ld = codes, codes_kind, rows, cols, ld
        tag.scale = 1.0
        tag.row_sum, tag.row_scale, tag.bits, tag.ld_bits = row_sum, row_scale, bits, ldb
        tag.overflow, tag.shape, tag.version, tag.layout = overflow, shape, None, layout
        tag.row_parts, tag.row_mul = row_parts, (1.0 / max(cols, 1) if row_parts else 1.0)
        tag.range_ok = overflow is None or (pre is not None and mode == L.Q_DOREFA
                                            and clamp_guarantees_lane(codes_kind, bit_width, pre[2], pre[3]))
        if _strict is True:
            tag.check()
    return y, tag


def codes_buffer(rows, cols, codes_kind, dev):
    """Uninitialised [rows, ld] code matrix as quant_act lays it out (8-bit or e2m1 lanes), for calls that fill row ranges of
    it (codes_out)."""
    if codes_kind == L.CODES_F4:
        return torch.empty((rows, round_up(max(cols, 1), 32) // 2), dtype=torch.uint8, device=dev)
    if codes_kind in (L.CODES_I8, L.CODES_U8):
        return torch.empty((rows, round_up(max(cols, 1), 16)), dtype=torch.int8 if codes_kind == L.CODES_I8 else torch.uint8,
                           device=dev)
    raise ValueError("codes_buffer: 8-bit and e2m1 lanes only")


class WeightPack:
    """k-bit weight matrix resident in HBM (the persistent format) + what the epilogue needs."""
    __slots__ = ("kind", "bit_width", "n", "k", "packed", "ld_packed", "alpha", "alpha_norm", "alpha_max", "stats",
                 "col_scale", "planes", "ld_planes", "wq", "wscale", "emin", "_prefetch", "_last_kind", "_first", "_hold",
                 "_hold_on", "_padw")
    # kind 'lin' / 'log' (LogLin layers): packed = int8 codes [1, n, ld]; value = code * wscale (lin) or
    # sign(code) * 2^(emin + |code| - 1) (log)

    def __init__(self):
        self.wscale, self.emin = 1.0, 0
        self._prefetch = None        # (out_kind, operand, ld, event): expanded ahead of time on a side stream
        self._last_kind = None       # operand kind the last contraction asked for (what a prefetch expands)
        self._first = None           # cached plane-pixel weights of a first conv layer (engine._first_layer_weights)
        self._padw = None            # cached channel-padded conv operand (engine._channel_padded_weights)
        self._hold, self._hold_on = None, False   # one expansion shared by the row bands of engine.linear_banded

    def nbytes(self):
        t = self.packed if self.packed is not None else self.planes
        return t.numel() * t.element_size()


def _lane_bits(k):
    return 1 if k <= 1 else 2 if k <= 2 else 4 if k <= 4 else 8


def pack_weight(w2d, kind, bit_width=1, want_wq=False, alpha=None):
    """fp32 [n, k] master weights -> packed k-bit HBM format (one reduction pass + one pack pass)."""
    require_cuda(w2d, "weight")
    w2d = as_f32c(w2d)
    n, k = w2d.shape
    dev = w2d.device
    mode = {"sign": L.W_SIGN, "ternary": L.W_TERNARY, "dorefa": L.W_DOREFA, "xnor": L.W_XNOR}[kind]
    lane = _lane_bits(bit_width) if kind == "dorefa" else 1
    ld_packed = round_up((k * lane + 7) // 8, 16)          # bytes, 16-byte rows
    planes = 2 if kind in ("ternary", "xnor") else 1
    p = WeightPack()
    p.kind, p.bit_width, p.n, p.k = kind, bit_width, n, k
    p.packed = torch.empty((planes, n, ld_packed), dtype=torch.uint8, device=dev)
    p.ld_packed = ld_packed
    p.stats = torch.empty(16, dtype=torch.float32, device=dev)
    p.alpha = None
    if kind == "xnor":
        p.alpha = alpha if alpha is not None else torch.empty(k, dtype=torch.float32, device=dev)
    p.wq = torch.empty_like(w2d) if want_wq else None
    p.planes, p.ld_planes, p.col_scale = None, 0, None
    p.alpha_norm = p.alpha_max = None
    a = L.QtWeightPack()
    a.mode, a.bit_width = mode, bit_width
    a.w, a.n, a.k, a.ld_w = _p(w2d), n, k, k
    a.packed, a.ld_packed = _p(p.packed), ld_packed
    a.alpha, a.alpha_is_input = _p(p.alpha), 1 if alpha is not None else 0
    a.stats, a.wq = _p(p.stats), _p(p.wq)
    L.check(L.lib().qt_pack_weight(C.byref(a), _stream()), "qt_pack_weight")
    if kind == "xnor":
        # fp16 fast route: alpha normalised to max 1 (fp16 normal range); the max goes back in through col_scale
        amax = torch.clamp_min(p.alpha.max(), 1e-30)
        p.alpha_norm = (p.alpha / amax).contiguous()
        p.alpha_max = amax.reshape(1).expand(n).contiguous()
    if kind == "dorefa":
        # per-output-column scale kept on the device (no host sync):
        #   k == 1 : E = mean|W|            (dorefa_connect.py:100-102)
        #   k >= 2 : 1/n, or 0 when W is all zeros (dorefa_connect.py:106-107)
        if bit_width == 1:
            p.col_scale = p.stats[1:2].expand(n).contiguous()
        else:
            inv_n = 1.0 / float(2 ** bit_width - 1)
            p.col_scale = torch.where(p.stats[3:4] == 0, torch.zeros_like(p.stats[3:4]),
                                      torch.full_like(p.stats[3:4], inv_n)).expand(n).contiguous()
    return p


def pack_real_weight(wq2d):
    """Arbitrary fp32 weights -> two bf16 planes (hi, lo) [2, n, ld]: the 'real' weight operand
    (LogLin quantized values are exact in the hi plane)."""
    require_cuda(wq2d, "weight")
    _, tag = quant_act(as_f32c(wq2d), L.Q_SPLIT, want_y=False, codes_kind=L.CODES_BF16X2, kind="real")
    p = WeightPack()
    p.kind, p.bit_width, p.n, p.k = "real", 32, wq2d.shape[0], wq2d.shape[1]
    p.packed, p.ld_packed, p.alpha, p.stats, p.col_scale, p.wq = None, 0, None, None, None, None
    p.alpha_norm = p.alpha_max = None
    p.planes, p.ld_planes = tag.codes, tag.ld
    return p


def pack_loglin_weight(wq2d, dtype, fsr, bit_width):
    """Lin / Log *quantized* weights (the values the reference's weight op returns) -> int8 codes [1, n, ld] (8 bits per weight
    in HBM instead of 32): lin code = wq / step, log code = sign * (log2|wq| - emin + 1).  bit_width <= 6."""
    require_cuda(wq2d, "weight")
    wq2d = as_f32c(wq2d)
    n, k = wq2d.shape
    ld = round_up(max(k, 1), 16)
    p = WeightPack()
    p.kind, p.bit_width, p.n, p.k = dtype, bit_width, n, k
    p.alpha = p.alpha_norm = p.alpha_max = p.stats = p.col_scale = p.planes = p.wq = None
    p.ld_planes = 0
    codes = torch.zeros((1, n, ld), dtype=torch.int8, device=wq2d.device)
    if dtype == "lin":
        p.wscale, p.emin = float(2.0 ** (fsr - bit_width)), 0
        codes[0, :, :k] = torch.round(wq2d / p.wscale).to(torch.int8)
    else:
        p.wscale, p.emin = 1.0, int(fsr - 2 ** bit_width)
        mag = torch.where(wq2d == 0, torch.zeros_like(wq2d), torch.log2(wq2d.abs().clamp_min(1e-38)) - p.emin + 1)
        codes[0, :, :k] = (torch.sign(wq2d) * torch.round(mag)).to(torch.int8)
    p.packed, p.ld_packed = codes, ld
    return p


def col_absmean(w2d):
    require_cuda(w2d, "weight")
    w2d = as_f32c(w2d)
    n, k = w2d.shape
    out = torch.empty(k, dtype=torch.float32, device=w2d.device)
    L.check(L.lib().qt_col_absmean(_p(w2d), n, k, k, _p(out), _stream()), "qt_col_absmean")
    return out


def expand_weight(p, out_kind):
    """Packed k-bit weights -> transient tensor-core operand (lives in L2 between the two kernels)."""
    p._last_kind = out_kind
    if p._hold_on and p._hold is not None and p._hold[0] == out_kind:
        return p._hold[1], p._hold[2]
    if p._hold_on:
        w, ld = expand_weight_now(p, out_kind)
        p._hold = (out_kind, w, ld)
        return w, ld
    return expand_weight_now(p, out_kind)


def expand_weight_now(p, out_kind):
    pf = p._prefetch
    if pf is not None:
        p._prefetch = None
        if pf[0] == out_kind:                    # expanded ahead of time (fusion.OperandPrefetch): wait for it, no launch here
            cur = torch.cuda.current_stream()
            cur.wait_event(pf[3])
            pf[1].record_stream(cur)
            return pf[1], pf[2]
    return _expand_weight(p, out_kind)


def _expand_weight(p, out_kind):
    dev = p.packed.device
    if p.kind in ("lin", "log"):
        if p.kind == "lin" and out_kind == L.CODES_I8:
            return p.packed[0], p.ld_packed                  # the HBM format IS the int8 operand
        if out_kind != L.CODES_BF16:
            raise RuntimeError("internal: LogLin weights expand to int8 (lin) or bf16 operands only")
        ld = round_up(p.k, 16)
        out = torch.empty((1, p.n, ld), dtype=torch.bfloat16, device=dev)
        L.check(L.lib().qt_expand_loglin(_p(p.packed), p.n, p.k, p.ld_packed, 1 if p.kind == "log" else 0, int(p.emin),
                                         _p(out), ld, _stream()), "qt_expand_loglin")
        return out, ld
    ld = round_up(p.k, 16)
    if out_kind == L.CODES_F4:
        ld = round_up(p.k, 32)
        out = torch.empty((p.n, ld // 2), dtype=torch.uint8, device=dev)
    elif out_kind == L.CODES_I8:
        out = torch.empty((p.n, ld), dtype=torch.int8, device=dev)
    elif out_kind == L.CODES_U8:
        out = torch.empty((p.n, ld), dtype=torch.uint8, device=dev)
    elif out_kind == L.CODES_BF16:
        out = torch.empty((1, p.n, ld), dtype=torch.bfloat16, device=dev)
    elif out_kind in (L.CODES_F16, L.CODES_F16_EXACT):
        out = torch.empty((1, p.n, ld), dtype=torch.float16, device=dev)
    else:
        out = torch.empty((2, p.n, ld), dtype=torch.bfloat16, device=dev)
    mode = {"sign": L.W_SIGN, "ternary": L.W_TERNARY, "dorefa": L.W_DOREFA, "xnor": L.W_XNOR}[p.kind]
    a = L.QtWeightExpand()
    a.mode, a.bit_width = mode, p.bit_width
    a.packed, a.n, a.k, a.ld_packed = _p(p.packed), p.n, p.k, p.ld_packed
    a.alpha = _p(p.alpha_norm if out_kind == L.CODES_F16 else p.alpha)
    a.out, a.out_kind, a.ld_out = _p(out), out_kind, ld
    L.check(L.lib().qt_expand_weight(C.byref(a), _stream()), "qt_expand_weight")
    return out, ld


def conv_weight_2d(w):
    """[O, C/g, kh, kw] conv weights -> [O, kh*kw*C/g] with the channel fastest: the K order of qt_im2col."""
    if w.dim() != 4:
        return w.reshape(w.shape[0], -1)
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def im2col(x4d, elem_bytes, geom, group, out, ld_out, row_sum=None, is_unsigned=False, nhwc=False, split3=False):
    """geom = (kh, kw, sh, sw, ph, pw, dh, dw, groups, OH, OW); x4d is [B,C,H,W] (nhwc=False) or [B,H,W,C]."""
    if nhwc:
        B, H, W, Cc = x4d.shape
    else:
        B, Cc, H, W = x4d.shape
    kh, kw, sh, sw, ph, pw, dh, dw, groups, OH, OW = geom
    a = L.QtIm2col()
    a.x, a.nhwc, a.elem_bytes, a.is_unsigned = _p(x4d), int(nhwc), elem_bytes, int(is_unsigned)
    a.B, a.C, a.H, a.W = B, Cc, H, W
    a.kh, a.kw, a.stride_h, a.stride_w, a.pad_h, a.pad_w, a.dil_h, a.dil_w = kh, kw, sh, sw, ph, pw, dh, dw
    a.groups, a.group, a.OH, a.OW = groups, group, OH, OW
    a.out, a.ld_out, a.row_sum, a.split3 = _p(out), ld_out, _p(row_sum), int(split3)
    L.check(L.lib().qt_im2col(C.byref(a), _stream()), "qt_im2col")


def make_epi(out, *, ldo, out_mode=0, nchw_inner=1, bias=None, row_scale=None, col_scale=None, row_sum=None,
             scale=1.0, acc_mul=1, rs_mul=0, acc_out=None, out_offset=0, row_parts=0, row_mul=1.0, requant=None,
             out_clamp=None, residual=None, ld_res=0, a_ready=None):
    """requant: a RequantOut (fused re-quantisation of the output); row_parts > 0: row_scale / row_sum are partial sums
    left by a previous layer's requant epilogue."""
    e = L.QtEpilogue()
    e.bias, e.row_scale, e.col_scale, e.row_sum = _p(bias), _p(row_scale), _p(col_scale), _p(row_sum)
    e.scale, e.acc_mul, e.rs_mul = float(scale), int(acc_mul), int(rs_mul)
    e.out = None if out is None else C.c_void_p(out.data_ptr() + 4 * out_offset)
    e.ldo, e.out_mode, e.nchw_inner = ldo, out_mode, nchw_inner
    e.acc_out = _p(acc_out)
    if row_parts > 0:
        if row_scale is not None:
            e.row_scale_parts, e.row_scale_mul = int(row_parts), float(row_mul)
        if row_sum is not None:
            e.row_sum_parts = int(row_parts)
    if out_clamp is not None:
        e.out_clamp, e.out_lo, e.out_hi = 1, float(out_clamp[0]), float(out_clamp[1])
    if requant is not None:
        e.requant = C.pointer(requant.c)
        e._keep = requant        # keep the ctypes struct alive as long as the epilogue
    if residual is not None:
        e.residual, e.ld_res = _p(residual), int(ld_res)
    if a_ready is not None:              # (int32 counters, rows per counter, count that means "block complete")
        e.a_ready, e.a_ready_rows, e.a_ready_target = _p(a_ready[0]), int(a_ready[1]), int(a_ready[2])
        e._keep_ready = a_ready[0]
    return e


class RequantOut:
    """Output operand of a fused requant epilogue (include/qtb200.h QtRequant): the next layer's low-bit codes,
    written by the tcgen05 epilogue instead of the fp32 activation.

    mode/bit_width/codes_kind as qt_quant_act; lo/hi: clamp in front of the quantizer (None: no clamp);
    rows x cols: logical shape of the codes ([M, N] of the producing GEMM); `codes` may be passed in (conv groups write
    column slices of one channels-last tensor)."""

    def __init__(self, mode, bit_width, codes_kind, rows, cols, dev, lo=None, hi=None, want_row_sum=False, codes=None,
                 ld=None, col_offset=0, cover=0):
        self.mode, self.bit_width, self.codes_kind, self.rows, self.cols = mode, bit_width, codes_kind, rows, cols
        self.ld = round_up(max(cols, 1), 32) if ld is None else ld
        if codes is None:
            if codes_kind == L.CODES_F4:
                codes = torch.empty((rows, self.ld // 2), dtype=torch.uint8, device=dev)
            elif codes_kind in (L.CODES_I8, L.CODES_U8):
                codes = torch.empty((rows, self.ld), dtype=torch.int8 if codes_kind == L.CODES_I8 else torch.uint8, device=dev)
            elif codes_kind == L.CODES_F16:
                codes = torch.empty((rows, self.ld), dtype=torch.float16, device=dev)
            elif codes_kind == L.CODES_BF16:
                codes = torch.empty((rows, self.ld), dtype=torch.bfloat16, device=dev)
            else:
                raise ValueError("RequantOut: unsupported codes_kind %r" % (codes_kind,))
        self.codes = codes
        cap = int(L.lib().qt_requant_max_parts(cols))
        self.row_part = torch.empty((cap, rows), dtype=torch.float32, device=dev) if mode == L.Q_XNOR_ROW else None
        self.row_sum_part = torch.empty((cap, rows), dtype=torch.int32, device=dev) if want_row_sum else None
        self.overflow = _overflow_flag(dev, mode, lo is not None and clamp_guarantees_lane(codes_kind, bit_width, lo, hi))
        c = L.QtRequant()
        c.mode, c.bit_width, c.codes_kind, c.ld_codes = mode, bit_width, codes_kind, self.ld
        elem_num, elem_den = {L.CODES_F4: (1, 2), L.CODES_I8: (1, 1), L.CODES_U8: (1, 1)}.get(codes_kind, (2, 1))
        c.codes = C.c_void_p(codes.data_ptr() + col_offset * elem_num // elem_den)
        c.clamp = 0 if lo is None else 1
        c.lo, c.hi = (0.0, 0.0) if lo is None else (float(lo), float(hi))
        c.row_part, c.row_sum_part, c.overflow = _p(self.row_part), _p(self.row_sum_part), _p(self.overflow)
        c.cover = int(cover)
        self.c = c

    @property
    def row_parts(self):
        return int(self.c.row_parts)


def is_channels_last(x):
    """4-D tensor whose memory is [B, H, W, C] dense (torch.channels_last), and not also plain-contiguous."""
    return (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
            and not (x.is_contiguous() and x.shape[1] != 1))


def quant_act_nhwc(x_cl, mode, *, bit_width=0, codes_kind=L.CODES_I8, kind=None):
    """Activation quantizer on a channels-last fp32 tensor [B, C, H, W] (memory [B, H, W, C]): the codes come out channels-last
    too, with no transpose -- the tensor is one [B*H*W, C] row matrix.  Code-only (no fp32 result).  C % 16 == 0."""
    require_cuda(x_cl, "input")
    B, Cn, Hn, Wn = x_cl.shape
    if not is_channels_last(x_cl) or x_cl.dtype != torch.float32 or Cn % 16:
        raise RuntimeError("internal: quant_act_nhwc needs a channels-last fp32 tensor with C % 16 == 0")
    rows = x_cl.permute(0, 2, 3, 1).reshape(B * Hn * Wn, Cn)          # a view: same memory
    _, tag = quant_act(rows, mode, bit_width=bit_width, want_y=False, codes_kind=codes_kind, kind=kind)
    tag.codes = tag.codes.view(B, Hn, Wn, Cn)
    tag.rows, tag.cols, tag.ld, tag.layout, tag.shape = B, Cn * Hn * Wn, Cn * Hn * Wn, "nhwc", (B, Cn, Hn, Wn)
    return tag


def conv_bf16(x_nhwc, geom_args, w, ldw, N, epi):
    """Implicit-GEMM conv on channels-last bf16 activations (qt_conv_bf16).  geom_args: the 16 QtConvGeom fields."""
    g = L.QtConvGeom(*geom_args)
    L.check(L.lib().qt_conv_bf16(_p(x_nhwc), C.byref(g), _p(w), ldw, N, C.byref(epi), _stream()), "qt_conv_bf16")


def image_planes(x, planes, pad_h, pad_w, Hp, Wp, fold_h=1, fold_w=1):
    """fp32 NCHW image -> zero-padded channels-last bf16 plane pixels (qt_image_planes): [B, Hp, Wp, 16], or with a
    space-to-depth fold [B, Hp / fold_h, Wp / fold_w, fold_h * fold_w * 16]."""
    require_cuda(x, "input")
    x = as_f32c(x)
    B, Cn, Hn, Wn = x.shape
    out = torch.empty((B, Hp // fold_h, Wp // fold_w, fold_h * fold_w * 16), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().qt_image_planes(_p(x), B, Cn, Hn, Wn, planes, pad_h, pad_w, Hp, Wp, fold_h, fold_w, _p(out), _stream()),
            "qt_image_planes")
    return out


def image_windows(x, planes, kw, stride_w, pad_h, pad_w, Hp, OW, slots):
    """fp32 NCHW image -> row-window records (qt_image_windows): [B, Hp, OW, slots] bf16, record (hp, ow) = the kw pixels
    w = ow * stride_w - pad_w + kx of row hp - pad_h, each as planes * C slots."""
    require_cuda(x, "input")
    x = as_f32c(x)
    B, Cn, Hn, Wn = x.shape
    out = torch.empty((B, Hp, OW, slots), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().qt_image_windows(_p(x), B, Cn, Hn, Wn, planes, kw, stride_w, pad_h, pad_w, Hp, OW, slots, _p(out), _stream()),
            "qt_image_windows")
    return out


def head_f32(x, weight, bias):
    """Global average pool + fp32 Linear (qt_head_f32), batch-invariant: x channels-last [B, C, H, W] (or [B, C]) fp32,
    weight [N, C], bias [N] or None -> [B, N]."""
    require_cuda(x, "input")
    if x.dim() == 4:
        if not is_channels_last(x) and not (x.shape[2] == 1 and x.shape[3] == 1):
            x = x.contiguous(memory_format=torch.channels_last)
        B, Cn, HW = x.shape[0], x.shape[1], x.shape[2] * x.shape[3]
    else:
        x = as_f32c(x)
        B, Cn, HW = x.shape[0], x.shape[1], 1
    if x.dtype != torch.float32:
        x = x.float()
    w = as_f32c(weight)
    b = None if bias is None else as_f32c(bias)
    out = torch.empty((B, w.shape[0]), dtype=torch.float32, device=x.device)
    L.check(L.lib().qt_head_f32(_p(x), B, HW, Cn, _p(w), w.stride(0), _p(b), w.shape[0], _p(out), out.stride(0), _stream()),
            "qt_head_f32")
    return out


def _pool_geom(B, H, W, Cn, k, s, p):
    (kh, kw), (sh, sw), (ph, pw) = k, s, p
    OH, OW = (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    return L.QtPoolGeom(B, H, W, Cn, kh, kw, sh, sw, ph, pw, OH, OW), OH, OW


def pool_codes(codes_nhwc, k, s, p, use_min=None):
    """Max-pool (per-channel min where use_min[c]) of channels-last 8-bit codes [B, H, W, C] (qt_pool_codes)."""
    B, H, W, Cn = codes_nhwc.shape
    g, OH, OW = _pool_geom(B, H, W, Cn, k, s, p)
    out = torch.empty((B, OH, OW, Cn), dtype=codes_nhwc.dtype, device=codes_nhwc.device)
    L.check(L.lib().qt_pool_codes(_p(codes_nhwc), int(codes_nhwc.dtype == torch.uint8), C.byref(g), _p(use_min), _p(out),
                                  _stream()), "qt_pool_codes")
    return out


def peer_push(src, dsts, ctas=32):
    """Copy the contiguous tensor `src` into every tensor of `dsts` (same size; peer-mapped buffers) with one kernel (qt_peer_push)."""
    require_cuda(src, "source")
    nbytes = src.numel() * src.element_size()
    arr = (C.c_void_p * len(dsts))(*[d.data_ptr() for d in dsts])
    L.check(L.lib().qt_peer_push(_p(src), arr, len(dsts), nbytes, ctas, _stream()), "qt_peer_push")


def rowsum_codes(codes2d):
    """int32 row sums of an 8-bit code matrix [rows, ld] (qt_rowsum_codes)."""
    rows, ld = codes2d.shape
    out = torch.empty(rows, dtype=torch.int32, device=codes2d.device)
    L.check(L.lib().qt_rowsum_codes(_p(codes2d), int(codes2d.dtype == torch.uint8), rows, ld, _p(out), _stream()), "qt_rowsum_codes")
    return out


def pool_quant_f32(x_cl, k, s, p, want_out=True, mode=None, bit_width=0, codes_kind=L.CODES_NONE):
    """Max-pool of a channels-last fp32 tensor fused with an activation quantizer (qt_pool_quant_f32).
    Returns (pooled channels-last fp32 tensor or None, codes [B, OH, OW, C] or None, overflow flag or None)."""
    B, Cn, H, W = x_cl.shape
    g, OH, OW = _pool_geom(B, H, W, Cn, k, s, p)
    out = torch.empty((B, Cn, OH, OW), dtype=torch.float32, device=x_cl.device, memory_format=torch.channels_last) if want_out else None
    codes = overflow = None
    if mode is not None:
        codes = torch.empty((B, OH, OW, Cn), dtype=torch.int8 if codes_kind == L.CODES_I8 else torch.uint8, device=x_cl.device)
        overflow = _overflow_flag(x_cl.device, mode)
    L.check(L.lib().qt_pool_quant_f32(_p(x_cl), C.byref(g), _p(out), -1 if mode is None else mode, bit_width, _p(codes),
                                      codes_kind, _p(overflow), _stream()), "qt_pool_quant_f32")
    return out, codes, overflow


def gemm_b1b1(a_bits, lda, w_bits, ldw, M, N, K, epi):
    L.check(L.lib().qt_gemm_b1b1(_p(a_bits), lda, _p(w_bits), ldw, M, N, K, C.byref(epi), _stream()), "qt_gemm_b1b1")


def gemm_b1t2(a_bits, lda, w_nz, w_sign, ldw, M, N, K, epi):
    L.check(L.lib().qt_gemm_b1t2(_p(a_bits), lda, _p(w_nz), _p(w_sign), ldw, M, N, K, C.byref(epi), _stream()),
            "qt_gemm_b1t2")


def gemm_i8(a, a_signed, lda, w, w_signed, ldw, M, N, K, epi, backend=L.BACKEND_AUTO):
    L.check(L.lib().qt_gemm_i8(_p(a), int(a_signed), lda, _p(w), int(w_signed), ldw, M, N, K, C.byref(epi),
                               backend, _stream()), "qt_gemm_i8")


def gemm_f4(a, lda, w, ldw, M, N, K, epi):
    """e2m1 codes x e2m1 codes on tcgen05 kind::mxf4 (unit block scales): exact integer accumulators."""
    L.check(L.lib().qt_gemm_f4(_p(a), lda, _p(w), ldw, M, N, K, C.byref(epi), _stream()), "qt_gemm_f4")


def gemm_f16(a, lda, a_plane_stride, w, ldw, w_plane_stride, passes, M, N, K, epi, backend=L.BACKEND_AUTO,
             fmt=L.FMT_BF16):
    n = len(passes)
    pa = (C.c_int * n)(*[p[0] for p in passes])
    pw = (C.c_int * n)(*[p[1] for p in passes])
    L.check(L.lib().qt_gemm_f16(_p(a), lda, a_plane_stride, _p(w), ldw, w_plane_stride, fmt, n, pa, pw, M, N, K,
                                C.byref(epi), backend, _stream()), "qt_gemm_f16")


def conv_i8(x_nhwc, a_signed, geom, group, w, w_signed, ldw, N, epi, corners=None):
    """Implicit-GEMM conv on channels-last codes; returns False when the shape needs the explicit im2col route.
    corners = (lower_w, upper_w): explicit bounding box of the window base positions along W (asymmetric padding)."""
    B, H, W, Cc = x_nhwc.shape
    kh, kw, sh, sw, ph, pw, dh, dw, groups, OH, OW = geom
    g = L.QtConvGeom(B, Cc, H, W, kh, kw, sh, sw, ph, pw, dh, dw, groups, group, OH, OW)
    if corners is not None:
        g.corner_mode, g.lower_w, g.upper_w = 1, int(corners[0]), int(corners[1])
    rc = L.lib().qt_conv_i8(_p(x_nhwc), int(a_signed), C.byref(g), _p(w), int(w_signed), ldw, N, C.byref(epi), _stream())
    if rc == -3:
        return False
    L.check(rc, "qt_conv_i8")
    return True


def patch_rowsum(x_nhwc, is_unsigned, geom, group):
    B, H, W, Cc = x_nhwc.shape
    kh, kw, sh, sw, ph, pw, dh, dw, groups, OH, OW = geom
    g = L.QtConvGeom(B, Cc, H, W, kh, kw, sh, sw, ph, pw, dh, dw, groups, group, OH, OW)
    S = torch.empty(B * H * W, dtype=torch.int32, device=x_nhwc.device)
    rs = torch.empty(B * OH * OW, dtype=torch.int32, device=x_nhwc.device)
    L.check(L.lib().qt_patch_rowsum(_p(x_nhwc), int(is_unsigned), C.byref(g), _p(S), _p(rs), _stream()), "qt_patch_rowsum")
    return rs


def gemm_f32(a, lda, w, ldw, M, N, K, epi):
    L.check(L.lib().qt_gemm_f32(_p(a), lda, _p(w), ldw, M, N, K, C.byref(epi), _stream()), "qt_gemm_f32")


def device_caps(device=0):
    maj, mnr, sms, tc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    L.check(L.lib().qt_device_caps(device, C.byref(maj), C.byref(mnr), C.byref(sms), C.byref(tc)), "qt_device_caps")
    return dict(sm_major=maj.value, sm_minor=mnr.value, num_sms=sms.value, has_tcgen05=bool(tc.value))


def transpose_split(x2d, planes=2):
    """fp32 [R, C] -> bf16 planes of the transpose [planes, C, ld] (ld = R rounded up to 8): the K-major operand of a
    gradient contraction whose reduction runs over the leading dimension of x2d."""
    require_cuda(x2d, "tensor")
    x2d = as_f32c(x2d)
    R, Cc = x2d.shape
    ld = round_up(max(R, 1), 8)
    out = torch.empty((planes, Cc, ld), dtype=torch.bfloat16, device=x2d.device)
    L.check(L.lib().qt_transpose_split(_p(x2d), R, Cc, Cc, _p(out), ld, planes, _stream()), "qt_transpose_split")
    return out, ld


def ste_clip(grad, x, thresh=1.001):
    """grad * 1[|x| <= thresh] in one pass (the clip-mask straight-through estimator)."""
    require_cuda(grad, "grad")
    g = as_f32c(grad)
    xx = as_f32c(x)
    out = torch.empty_like(g)
    L.check(L.lib().qt_ste_clip(_p(g), _p(xx), float(thresh), _p(out), g.numel(), _stream()), "qt_ste_clip")
    return out
