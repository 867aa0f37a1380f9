"""CPU oracle for the QuantTorch quantized forward hot path.

TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module; the product package (``pytorch_quantize_impls_b200``) never does and
raises when its CUDA library is missing.

What it is: a restatement, in plain CPU torch + numpy, of the reference's
algorithm for the path named by BASELINE.json -- fake-quant op followed by a
dense fp32 ``F.linear`` / ``F.conv2d``.  The reference itself is pure Python on
torch (no native code), so the restatement uses the same torch primitives in
the same order; that makes it bit-identical to the live reference on CPU
(checked in ``tests/test_oracle_vs_reference.py`` whenever ``/root/reference``
is present, and against the committed vectors in ``tests/golden/`` everywhere).

Parity pinning: PINNED for BinaryNet / Terner / DoReFa / LogLin by the
reference's own known-answer tests (ported in
``tests/test_oracle_golden.py::test_reference_kats``) and by golden vectors
generated from the live reference (``oracle/gen_golden.py``, ``gen_golden_grads.py``,
``gen_golden_loglin.py``, ``gen_golden_functional.py``).  XnorNet: the reference's own XNOR tests are empty
files (tests/implementations/XNOR/*.py, 0 bytes), so XnorNet parity is pinned
only by outputs of the reference run here (golden vectors), not by reference
KATs.

Citations are ``path:line`` relative to /root/reference/.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# elementwise quantizers
# --------------------------------------------------------------------------


def safe_sign(x: torch.Tensor) -> torch.Tensor:
    """QuantTorch/functions/common.py:4-7 -- sign with 0 -> +1 (also -0.0, NaN -> +1)."""
    r = torch.sign(x)
    r[r == 0] = 1
    return r


def binary_det(x: torch.Tensor) -> torch.Tensor:
    """QuantTorch/functions/binary_connect.py:22-28 (forward of BinaryConnectDeterministic)."""
    return safe_sign(x)


def ste_clip_mask(x: torch.Tensor) -> torch.Tensor:
    """binary_connect.py:30-38 / terner_connect.py:29-34: grad passes where |x| <= 1.001."""
    return (torch.abs(x) <= 1.001).to(x.dtype)


def ternary_det(x: torch.Tensor) -> torch.Tensor:
    """QuantTorch/functions/terner_connect.py:24-27.

    (s + safeSign(x - 0.5 s)) / 2 with s = safeSign(x):
    x >= 0.5 -> +1, -0.5 <= x < 0.5 -> 0, x < -0.5 -> -1 (ties: +0.5 -> 1, -0.5 -> 0).
    """
    s = safe_sign(x)
    return (s + safe_sign(x - 0.5 * s)) / 2


def dorefa_quantize(x: torch.Tensor, bit_width: int = 3) -> torch.Tensor:
    """QuantTorch/functions/dorefa_connect.py:11-25 (_quantize).

    k == 1 -> safeSign, k == 32 -> identity, else fl(1/(2^k-1)) * round((2^k-1) x);
    no clamp, torch.round = half-to-even.
    """
    if bit_width == 1:
        return safe_sign(x)
    if bit_width == 32:
        return x
    two = torch.ones_like(x) * 2
    n = torch.pow(two, bit_width) - 1
    return (1 / n) * torch.round(n * x)


def dorefa_weight(w: torch.Tensor, bit_width: int = 3) -> torch.Tensor:
    """QuantTorch/functions/dorefa_connect.py:99-111 (nnQuantWeight._QuantWeight.forward)."""
    if bit_width == 1:
        e = torch.mean(torch.abs(w))
        return dorefa_quantize(w, 1) * e          # _ignore_factor_op :66-79
    if bit_width == 32:
        return w
    if torch.max(torch.abs(w)) == 0.0:            # :106-107
        return torch.zeros_like(w)
    t = torch.tanh(w)
    t = t / (2 * torch.max(torch.abs(t))) + 0.5
    return 2 * dorefa_quantize(t, bit_width) - 1


def xnor_act(x: torch.Tensor, dim: int = 1) -> torch.Tensor:
    """QuantTorch/functions/xnor_connect.py:20-28 (_QuantXNOR.forward).

    sign(x) * mean(x, dim) -- plain mean (not mean|x|), torch.sign (0 -> 0).
    dim in {-1 (scalar), 0 (per column), 1 (per row)}.
    """
    if dim not in (-1, 0, 1):
        raise RuntimeError("dim must be -1, 0 or 1")        # xnor_connect.py:41-42
    if dim < 0:
        return torch.sign(x) * torch.mean(x)
    m = torch.mean(x, dim)
    return torch.sign(x) * m.view({0: (1, -1), 1: (-1, 1)}[dim])


def xnor_weight(w: torch.Tensor) -> torch.Tensor:
    """QuantTorch/functions/xnor_connect.py:111-112: sign(W) * mean(|W|, DIM=0, keepdim) -> alpha is [1, K]."""
    a = torch.mean(torch.abs(w), 0, keepdim=True)
    return torch.sign(w) * a


def xnor_conv_weight(w: torch.Tensor, dim=(0, 1)) -> torch.Tensor:
    """QuantTorch/functions/xnor_connect.py:140-141: sign(W) * mean(|W|, dim=[0,1], keepdim)."""
    a = torch.mean(torch.abs(w), list(dim), keepdim=True)
    return torch.sign(w) * a


def log_quant(x: torch.Tensor, fsr: int = 7, bit_width: int = 3, with_sign: bool = True) -> torch.Tensor:
    """QuantTorch/functions/log_lin_connect.py:29-32 (LogQuant forward)."""
    p = torch.pow(torch.ones_like(x) * 2,
                  torch.clamp(torch.round(torch.log2(torch.abs(x))), fsr - 2 ** bit_width, fsr))
    return torch.sign(x) * p if with_sign else p


def lin_quant(x: torch.Tensor, fsr: int = 7, bit_width: int = 3, with_sign: bool = True) -> torch.Tensor:
    """QuantTorch/functions/log_lin_connect.py:61-68 (LinQuant forward)."""
    if bit_width == 32:
        return x
    step = torch.FloatTensor([2]).pow(fsr - bit_width)
    if with_sign:
        return torch.sign(x) * torch.clamp(torch.round(torch.abs(x) / step) * step, 0, 2 ** fsr)
    return torch.clamp(torch.round(x / step) * step, 0, 2 ** fsr)


def loglin_weight(w, dtype="lin", fsr=7, bit_width=3):
    """QuantTorch/functions/log_lin_connect.py:84-100 (nnQuant dispatch, with_sign=True as the layers use it)."""
    if dtype == "lin":
        return lin_quant(w, fsr, bit_width, True)
    if dtype == "log":
        return log_quant(w, fsr, bit_width, True)
    raise RuntimeError("Only 'log' and 'lin' dtype are supported !")


# --------------------------------------------------------------------------
# layer forwards (training-mode forward: fake-quant the weight every call)
# --------------------------------------------------------------------------

def _conv(x, wq, b, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(x, wq, b, stride, padding, dilation, groups)


def linear_bin(x, w, b=None):
    """QuantTorch/layers/binary_layers.py:42-46."""
    return F.linear(x, binary_det(w), b)


def conv_bin(x, w, b=None, **kw):
    """QuantTorch/layers/binary_layers.py:103-106."""
    return _conv(x, binary_det(w), b, **kw)


def linear_ter(x, w, b=None):
    """QuantTorch/layers/terner_layers.py:47-51."""
    return F.linear(x, ternary_det(w), b)


def conv_ter(x, w, b=None, **kw):
    """QuantTorch/layers/terner_layers.py:89-92."""
    return _conv(x, ternary_det(w), b, **kw)


def linear_dorefa(x, w, b=None, bit_width=3):
    """QuantTorch/layers/dorefa_layers.py:41-45."""
    return F.linear(x, dorefa_weight(w, bit_width), b)


def conv_dorefa(x, w, b=None, bit_width=3, **kw):
    """QuantTorch/layers/dorefa_layers.py:77-82."""
    return _conv(x, dorefa_weight(w, bit_width), b, **kw)


def linear_xnor(x, w, b=None):
    """QuantTorch/layers/xnor_layers.py:33-34 -> xnor_connect.py:110-116."""
    return F.linear(x, xnor_weight(w), b)


def conv_xnor(x, w, b=None, dim=(0, 1), **kw):
    """QuantTorch/layers/xnor_layers.py:67-69 -> xnor_connect.py:139-146 (quant_input never enabled, xnor_layers.py:49)."""
    return _conv(x, xnor_conv_weight(w, dim), b, **kw)


def linear_loglin(x, w, b=None, dtype="lin", fsr=7, bit_width=3):
    """QuantTorch/layers/log_lin_layers.py:40-42."""
    return F.linear(x, loglin_weight(w, dtype, fsr, bit_width), b)


def conv_loglin(x, w, b=None, dtype="lin", fsr=7, bit_width=3, **kw):
    """QuantTorch/layers/log_lin_layers.py:87-93 (training-mode branch)."""
    return _conv(x, loglin_weight(w, dtype, fsr, bit_width), b, **kw)


# --------------------------------------------------------------------------
# functional dense / conv ops with hand-written backward
# Each returns (output, weight_q); `*_grads` restate the reference's backward formulas.
# --------------------------------------------------------------------------

def ternary_functional_weight(w: torch.Tensor) -> torch.Tensor:
    """QuantTorch/functions/terner_connect.py:83-90 (deterministic branch): built on torch.sign, NOT safeSign, so
    0 -> 0, +0.5 -> +0.5 and -0.5 -> -0.5 -- this differs from the layer path (ternary_det)."""
    sign = torch.sign(w)
    return (sign + torch.sign(w - 0.5 * sign)) / 2


def dorefa_functional_weight(w: torch.Tensor, bit_width: int, conv: bool) -> torch.Tensor:
    """QuantDense: dorefa_connect.py:127-133 (normalises by max|tanh W|); QuantConv2d: :166-172 (by tanh(max|W|), the same
    value up to rounding).  No all-zero guard here (unlike nnQuantWeight)."""
    if bit_width == 1:
        return safe_sign(w) * torch.mean(torch.abs(w)).detach()
    if bit_width == 32:
        return w
    m = torch.tanh(torch.max(torch.abs(w))) if conv else torch.max(torch.abs(torch.tanh(w)))
    return 2 * dorefa_quantize(0.5 + torch.tanh(w) / (2 * m), bit_width) - 1


def binary_dense(x, w, b=None):
    """QuantTorch/functions/binary_connect.py:96-101."""
    wq = safe_sign(w)
    return F.linear(x, wq, b), wq


def ternary_dense(x, w, b=None):
    """QuantTorch/functions/terner_connect.py:83-93."""
    wq = ternary_functional_weight(w)
    return F.linear(x, wq, b), wq


def ternary_conv2d(x, w, b=None, **kw):
    """QuantTorch/functions/terner_connect.py:122-131."""
    wq = ternary_functional_weight(w)
    return _conv(x, wq, b, **kw), wq


def quant_dense(x, w, b=None, bit_width=3):
    """QuantTorch/functions/dorefa_connect.py:125-136."""
    wq = dorefa_functional_weight(w, bit_width, conv=False)
    return F.linear(x, wq, b), wq


def quant_conv2d(x, w, b=None, bit_width=3, **kw):
    """QuantTorch/functions/dorefa_connect.py:164-175."""
    wq = dorefa_functional_weight(w, bit_width, conv=True)
    return _conv(x, wq, b, **kw), wq


def dense_grads(go, x, wq, has_bias=True):
    """grad_input = g . W_q, grad_weight = g^T . x (no STE mask), grad_bias = sum g: binary_connect.py:103-112,
    terner_connect.py:95-105, dorefa_connect.py:138-152 (before the tanh factor)."""
    return go.mm(wq), go.t().mm(x), (go.sum(0) if has_bias else None)


def conv_grads(go, x, w_shape, wq, has_bias=True, **kw):
    """terner_connect.py:133-149 / dorefa_connect.py:177-196 (before the tanh factor)."""
    gi = torch.nn.grad.conv2d_input(x.shape, wq, go, **kw)
    gw = torch.nn.grad.conv2d_weight(x, w_shape, go, **kw)
    # bias gradient: the reference reduces one axis at a time (terner_connect.py:144, dorefa_connect.py:190), which fixes
    # the fp32 summation order
    gb = go.sum(0).squeeze(0).sum(1).squeeze(1).sum(-1).squeeze(-1) if has_bias else None
    return gi, gw, gb


def dorefa_functional_grad_weight(gw, w, bit_width, conv):
    """The tanh chain of the DoReFa functional ops: dorefa_connect.py:146-149 (dense) / :186-187 (conv)."""
    if bit_width in (1, 32):
        return gw
    m = torch.tanh(torch.max(torch.abs(w))) if conv else torch.max(torch.abs(torch.tanh(w)))
    return gw * (1 - torch.pow(torch.tanh(w), 2)) / m


# --------------------------------------------------------------------------
# integer oracles (exact accumulators, int64)
# --------------------------------------------------------------------------

def sign_codes(x: torch.Tensor) -> np.ndarray:
    """+1/-1 int64 codes of safe_sign(x)  (bit = !(x < 0))."""
    return safe_sign(x).to(torch.int64).numpy()


def sign_bits_packed(x: torch.Tensor) -> np.ndarray:
    """Bit-pack rows of a 2-D tensor: bit i of uint32 word j <-> column 32 j + i, 1 <-> +1.
    Columns are padded to a multiple of 32 with 0 bits."""
    b = (~(x < 0)).numpy().astype(np.uint8)
    rows, cols = b.shape
    words = (cols + 31) // 32
    pad = np.zeros((rows, words * 32), np.uint8)
    pad[:, :cols] = b
    return np.packbits(pad.reshape(rows, words, 32), axis=-1, bitorder="little").view(np.uint32).reshape(rows, words)


def ternary_codes(x: torch.Tensor) -> np.ndarray:
    return ternary_det(x).to(torch.int64).numpy()


def dorefa_act_codes(x: torch.Tensor, k: int) -> np.ndarray:
    """Integer code c = round((2^k-1) x) (half-even, unclamped) so that dorefa_quantize(x,k) = fl(1/n) * c."""
    if k == 1:
        return sign_codes(x)
    n = float(2 ** k - 1)
    return torch.round(n * x).to(torch.int64).numpy()


def dorefa_weight_codes(w: torch.Tensor, k: int) -> np.ndarray:
    """Integer code c in [0, n] with W_q = (2 c - n) / n  (k in 2..8); k == 1 -> +-1 codes (scale E separately)."""
    if k == 1:
        return sign_codes(w)
    n = float(2 ** k - 1)
    if torch.max(torch.abs(w)) == 0.0:
        return None
    t = torch.tanh(w)
    t = t / (2 * torch.max(torch.abs(t))) + 0.5
    return torch.round(n * t).to(torch.int64).numpy()


def e2m1_pack(codes: np.ndarray, ld: int) -> np.ndarray:
    """Integer codes in [-4, 4] -> fp4 (e2m1) nibbles, two per byte (element 2j in the low nibble), rows zero-padded to
    `ld` elements.  e2m1 magnitudes: 0, 0.5, 1, 1.5, 2, 3, 4, 6 <-> 0..7, sign in bit 3 -- the operand format of the
    tcgen05 kind::mxf4 contraction (the reference has no packed format: binary_layers.py:42-46 feeds fp32 +-1)."""
    mag = np.array([0, 2, 4, 5, 6], np.uint8)
    c = np.asarray(codes, np.int64)
    assert np.abs(c).max(initial=0) <= 4
    rows, cols = c.shape
    nib = np.zeros((rows, ld), np.uint8)
    nib[:, :cols] = mag[np.abs(c)] | ((c < 0).astype(np.uint8) << 3)
    return (nib[:, 0::2] | (nib[:, 1::2] << 4)).astype(np.uint8)


def int_acc(codes_a: np.ndarray, codes_w: np.ndarray) -> np.ndarray:
    """Exact accumulator: A[M,K] . W[N,K]^T in int64."""
    return codes_a.astype(np.int64) @ codes_w.astype(np.int64).T


def xnor_popcount_acc(a_bits: np.ndarray, w_bits: np.ndarray, k: int) -> np.ndarray:
    """K - 2 popc(a ^ w) on packed uint32 rows (the identity the 1-bit kernel implements)."""
    x = a_bits[:, None, :] ^ w_bits[None, :, :]
    pc = np.unpackbits(x.view(np.uint8), axis=-1).sum(-1).astype(np.int64)
    return k - 2 * pc


# --------------------------------------------------------------------------
# model-level forward of BASELINE config[1] (XnorNet MLP), used by bench cpu legs
# --------------------------------------------------------------------------

def xnor_mlp_forward(x, weights, biases):
    """nnQuantXnor(1) -> LinearXNOR, three times (4096-4096-4096-1000 in the bench)."""
    h = x
    for w, b in zip(weights, biases):
        h = linear_xnor(xnor_act(h, 1), w, b)
    return h


def binary_mlp_layer(x, w, b):
    """BinaryConnect() -> LinearBin (north-star shape)."""
    return linear_bin(binary_det(x), w, b)
