"""GPU: the LogLin k-bit format.  Lin / Log weights live as int8 codes in HBM, the Lin activation quantizer
emits int8 codes (value / step) and the Log quantizer bf16 powers of two, so that
  * Lin x Lin is an exact integer contraction on tcgen05 kind::i8 (accumulators equal the int64 oracle, output equals the
    reference's fp32 path up to its own rounding),
  * Log weights / activations are exact single bf16 planes (one tensor pass instead of five).
Reference: QuantTorch/functions/log_lin_connect.py:9-80, layers/log_lin_layers.py:6-93 (restated in oracle/)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import quanttorch_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def Q():
    import pytorch_quantize_impls_b200 as Q
    return Q


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("dtype,fsr,bw,with_sign", [("lin", 2, 4, True), ("lin", 0, 3, False), ("lin", 7, 3, True), ("log", 2, 3, True),
                                                  ("log", 7, 2, True)])
def test_activation_quantizer_codes(Q, dtype, fsr, bw, with_sign):
    torch.manual_seed(5)
    x = torch.randn(37, 200) * (2.0 ** fsr) * 0.6
    x[0, :6] = torch.tensor([0.0, -0.0, 2.0 ** fsr, -(2.0 ** fsr) * 3, 2.0 ** (fsr - bw) * 0.5, 1e-30])
    ref = (O.lin_quant if dtype == "lin" else O.log_quant)(x, fsr, bw, with_sign)
    y = Q.functions.Quant(x.cuda(), dtype=dtype, fsr=fsr, bit_width=bw, with_sign=with_sign)
    assert torch.equal(y.cpu(), ref)
    tag = y._qt_codes
    assert tag is not None and tag.kind == dtype
    codes = tag.codes[:, :200].float().cpu()
    if dtype == "lin":
        step = 2.0 ** (fsr - bw)
        assert tag.scale == step and torch.equal(codes * step, ref)
        assert torch.equal(tag.row_sum.cpu(), codes.sum(1).to(torch.int32))
    else:
        assert torch.equal(codes, ref)                       # powers of two are exact in bf16


@pytest.mark.parametrize("fsr_w,bw_w,fsr_a,bw_a", [(0, 4, 2, 4), (-1, 3, 1, 6), (2, 6, 0, 2)])
def test_lin_times_lin_is_an_exact_integer_contraction(Q, fsr_w, bw_w, fsr_a, bw_a):
    torch.manual_seed(9)
    M, K, N = 130, 520, 264
    lay = Q.layers.LinearQuant(K, N, dtype="lin", fsr=fsr_w, bit_width=bw_w)
    lay.bias.data.uniform_(-1, 1)
    x = torch.randn(M, K) * (2.0 ** fsr_a) * 0.5
    xq = O.lin_quant(x, fsr_a, bw_a, True)
    wq = O.lin_quant(lay.weight.data, fsr_w, bw_w, True)
    ref = O.linear_loglin(xq, lay.weight.data, lay.bias.data, "lin", fsr_w, bw_w)
    # exact value of every output: integer accumulator * step_a * step_w (+ bias)
    sa, sw = 2.0 ** (fsr_a - bw_a), 2.0 ** (fsr_w - bw_w)
    acc = (xq / sa).to(torch.int64) @ (wq / sw).to(torch.int64).t()
    exact = acc.double() * sa * sw + lay.bias.data.double()
    lay = lay.cuda().eval()
    pack = lay._current_pack()
    assert pack.kind == "lin" and pack.packed.dtype == torch.int8 and pack.packed.shape[1] == N      # 8 bits per weight in HBM
    with torch.no_grad():
        y = lay(Q.functions.Quant(x.cuda(), "lin", fsr_a, bw_a, True))
        with Q.code_only_activations():
            y2 = lay(Q.functions.Quant(x.cuda(), "lin", fsr_a, bw_a, True))
    assert torch.equal(y, y2)
    assert float((y.double().cpu() - exact).abs().max()) <= 1e-6 * float(exact.abs().max())          # one fp32 rounding
    assert rel(y, ref) < 1e-5


@pytest.mark.parametrize("dtype,fsr,bw", [("log", 1, 3), ("log", 0, 2), ("lin", 1, 3)])
def test_loglin_layers_with_real_and_log_activations(Q, dtype, fsr, bw):
    torch.manual_seed(11)
    M, K, N = 96, 300, 72
    lay = Q.layers.LinearQuant(K, N, dtype=dtype, fsr=fsr, bit_width=bw)
    lay.bias.data.uniform_(-1, 1)
    x = torch.randn(M, K)
    ref_real = O.linear_loglin(x, lay.weight.data, lay.bias.data, dtype, fsr, bw)
    xl = O.log_quant(x, 1, 3, True)
    ref_log = O.linear_loglin(xl, lay.weight.data, lay.bias.data, dtype, fsr, bw)
    conv = Q.layers.QuantConv2d(32, 40, 3, padding=1, fsr=fsr, bit_width=bw, dtype=dtype)
    xi = torch.randn(3, 32, 9, 9)
    ref_conv = O.conv_loglin(O.lin_quant(xi, 1, 4, True), conv.weight.data, conv.bias.data, dtype, fsr, bw, padding=1)
    lay, conv = lay.cuda(), conv.cuda()          # training mode: the reference quantizes the conv weights in train mode only
    with torch.no_grad():
        assert rel(lay(x.cuda()), ref_real) < 5e-5
        assert rel(lay(Q.functions.Quant(x.cuda(), "log", 1, 3, True)), ref_log) < 1e-5
        assert rel(conv(Q.functions.Quant(xi.cuda(), "lin", 1, 4, True)), ref_conv) < 1e-5
    assert lay._make_pack(lay.weight).kind == dtype


def _golden_loglin_cases():
    import os, re
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "quanttorch_ref_loglin_v1.npz"))
    cases = {}
    for k in z.files:
        name, field = k.split("/")
        cases.setdefault(name, {})[field] = torch.from_numpy(z[k].copy()) if z[k].ndim else int(z[k])
    return cases


@pytest.mark.parametrize("name", sorted(_golden_loglin_cases()))
def test_loglin_chains_match_live_reference_goldens(Q, name):
    """LogLin layers fed by Lin / Log quantized activations vs outputs of the LIVE reference (oracle/gen_golden_loglin.py):
    the activation quantizer is bit-exact, the layer output within 1e-5 (Lin x Lin: exact integers, one fp32 rounding)."""
    import re
    c = _golden_loglin_cases()[name]
    kind, da, fa, ba, dw, fw, bw = re.fullmatch(r"(dense|conv)_(lin|log)(-?\d+)_(\d+)__(lin|log)(-?\d+)_(\d+)", name).groups()
    fa, ba, fw, bw = int(fa), int(ba), int(fw), int(bw)
    if kind == "dense":
        lay = Q.layers.LinearQuant(c["w"].shape[1], c["w"].shape[0], dtype=dw, fsr=fw, bit_width=bw)
    else:
        lay = Q.layers.QuantConv2d(c["w"].shape[1], c["w"].shape[0], 3, stride=c["stride"], padding=c["padding"], fsr=fw,
                                   bit_width=bw, dtype=dw)
    lay.weight.data.copy_(c["w"]); lay.bias.data.copy_(c["b"])
    lay = lay.cuda()           # training mode, as the golden conv cases were produced
    with torch.no_grad():
        xq = Q.functions.Quant(c["x"].cuda(), da, fa, ba, True)
        assert torch.equal(xq.cpu(), c["xq"])
        assert rel(lay(xq), c["out"]) < 1e-5
