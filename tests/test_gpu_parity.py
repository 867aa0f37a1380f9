"""GPU parity tests proper: the CUDA path (through the C ABI) against the golden vectors of the live
reference and against the CPU oracle on seeded inputs.

Tolerances (north star): integer / code / accumulator work is BIT-EXACT; floating point after scaling
is compared as max|y - y_ref| / max|y_ref| <= 1e-3 (most cases are far tighter and say so)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import quanttorch_oracle as O  # noqa: E402
from conftest import case  # noqa: E402

REL = 1e-3


def relerr(y, ref):
    ref = ref.double()
    return float((y.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def Q():
    import pytorch_quantize_impls_b200 as Q
    assert torch.cuda.is_available()
    return Q


def cu(t):
    return t.cuda()


def eqnan(a, b):
    return torch.equal(torch.nan_to_num(a.cpu(), nan=777.0), torch.nan_to_num(b, nan=777.0))


# ------------------------------------------------------------------ elementwise ops: bit exact
def test_elementwise_ops_bit_exact(Q, golden):
    F = Q.functions
    c = case(golden, "safe_sign"); assert eqnan(F.safeSign(cu(c["x"])), c["out"])
    c = case(golden, "binary_det"); assert eqnan(F.BinaryConnectDeterministic.apply(cu(c["x"])), c["out"])
    c = case(golden, "ternary_det"); assert eqnan(F.TernaryConnectDeterministic.apply(cu(c["x"])), c["out"])
    c = case(golden, "ternary_kat1"); assert F.TernaryConnect()(cu(c["x"])).cpu().tolist() == [1, 1, 0, 0, -1, 0]
    c = case(golden, "ternary_kat2")
    assert F.TernaryConnectDeterministic.apply(cu(c["x"])).cpu().tolist() == [1, 0, 1, 0, 0, -1, 0, 1]
    for k in (1, 2, 3, 4, 5, 8, 32):
        c = case(golden, f"dorefa_quant_k{k}")
        assert eqnan(F.DorefaQuant(cu(c["x"]), k), c["out"]), k
        assert eqnan(F.nnDorefaQuant(k)(cu(c["x"])), c["out"]), k
    for d in (-1, 0, 1):
        c = case(golden, f"xnor_act_dim{d}")
        y = F.QuantXnor(cu(c["x"]), d).cpu()
        assert relerr(y, c["out"]) < 2e-6, d            # row mean: fp64-accumulated vs torch's fp32 tree sum
        assert torch.equal(torch.sign(y), torch.sign(c["out"]))
    for fsr, bw in ((7, 3), (2, 2), (5, 4)):
        c = case(golden, f"log_quant_{fsr}_{bw}"); assert eqnan(F.Quant(cu(c["x"]), "log", fsr, bw), c["out"])
        c = case(golden, f"lin_quant_{fsr}_{bw}"); assert eqnan(F.Quant(cu(c["x"]), "lin", fsr, bw), c["out"])
    c = case(golden, "lin_quant_unsigned_2_8")
    assert eqnan(F.Quant(cu(c["x"]), "lin", 2, 8, with_sign=False), c["out"])


def test_input_not_mutated_and_new_tensor(Q):
    x = torch.randn(5, 7).cuda()
    x0 = x.clone()
    y = Q.functions.safeSign(x)
    assert torch.equal(x, x0) and y.data_ptr() != x.data_ptr()


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8, 32])
def test_weight_quantizer(Q, golden, k):
    c = case(golden, f"dorefa_weight_k{k}")
    wq = Q.functions.nnQuantWeight(k)(cu(c["w"])).cpu()
    if k in (1, 32):
        assert relerr(wq, c["out"]) < 1e-6
    else:
        # codes are integers; device tanh (fp64, correctly rounded) vs CPU Sleef tanh may differ by 1 ulp, which can
        # move a value sitting exactly on a rounding boundary by one level: allow <= 0.1 % such elements, 1 level max
        n = 2 ** k - 1
        lv = torch.round((wq.double() + 1) * n / 2)
        lr = torch.round((c["out"].double() + 1) * n / 2)
        d = (lv - lr).abs()
        assert d.max() <= 1 and (d > 0).float().mean() <= 1e-3
        assert torch.allclose(wq, (2 * lv - n).float() / n, atol=3e-7)
    c = case(golden, "dorefa_weight_zero")
    assert torch.equal(Q.functions.nnQuantWeight(3)(cu(c["w"])).cpu(), c["out"])


# ------------------------------------------------------------------ dense layers vs golden
def _set(layer, w, b=None):
    layer.weight.data.copy_(w)
    if b is not None:
        layer.bias.data.copy_(b)
    return layer


def test_linear_bin_golden(Q, golden):
    c = case(golden, "lin_bin")
    K, N = c["w"].shape[1], c["w"].shape[0]
    lay = _set(Q.layers.LinearBin(K, N).cuda(), c["w"], c["b"])
    act = Q.functions.BinaryConnect()
    with torch.no_grad():
        xq = act(cu(c["x"]))
        assert torch.equal(xq.cpu(), c["xq"])
        y = lay(xq)                                   # tagged: 1-bit x 1-bit integer path
        assert torch.equal(y.cpu(), c["out"])         # exact integers + one bias rounding -> bit exact
        for kw in (dict(popcount=True), dict(popcount=False, i8="simt"), dict(i8="tcgen05")):
            Q.set_backend(**kw)
            assert torch.equal(lay(act(cu(c["x"]))).cpu(), c["out"]), kw
        Q.set_backend(i8="auto", popcount=False)
        y2 = lay(cu(c["xq"]))                          # untagged +-1 tensor: real-activation route
        assert relerr(y2, c["out"]) < 1e-5
        y3 = lay(cu(c["x"]))                           # arbitrary real input
        assert relerr(y3, c["out_real"]) < 5e-5
        # train/eval swap (BinaryNet/layer_test.py:96-127)
        w0 = lay.weight.data.clone()
        lay.train(False)
        c2 = case(golden, "lin_bin_eval")
        assert torch.equal(lay.weight.data.cpu(), c2["w_eval"])
        assert torch.equal(lay(act(cu(c["x"]))).cpu(), c2["out"])
        lay.train(True)
        assert torch.equal(lay.weight.data, w0)
    c = case(golden, "lin_bin_kat")
    lay = _set(Q.layers.LinearBin(3, 1, bias=False).cuda(), c["w"])
    assert lay(cu(c["x"])).item() == pytest.approx(6.0, abs=1e-4)


def test_linear_ter_golden(Q, golden):
    c = case(golden, "lin_ter")
    N, K = c["w"].shape
    lay = _set(Q.layers.LinearTer(K, N).cuda(), c["w"], c["b"])
    act = Q.functions.BinaryConnect()
    with torch.no_grad():
        for kw in (dict(popcount=True), dict(popcount=False, i8="simt"), dict(i8="tcgen05")):
            Q.set_backend(**kw)
            assert torch.equal(lay(act(cu(c["x"]))).cpu(), c["out"]), kw
        Q.set_backend(i8="auto", popcount=False)
        assert relerr(lay(cu(c["x"])), c["out_real"]) < 5e-5


@pytest.mark.parametrize("k", [1, 2, 3, 4, 8])
def test_linear_dorefa_golden(Q, golden, k):
    with torch.no_grad():
        for ka in (k, 8 if k != 8 else 4):
            c = case(golden, f"lin_dorefa_w{k}a{ka}")
            N, K = c["w"].shape
            lay = _set(Q.layers.LinearDorefa(K, N, bit_width=k).cuda(), c["w"], c["b"])
            xq = Q.functions.DorefaQuant(cu(c["x"]), ka)
            assert torch.equal(xq.cpu(), c["xq"])
            for be in ("simt", "tcgen05"):
                Q.set_backend(i8=be)
                assert relerr(lay(xq), c["out"]) < 2e-5, (k, ka, be)
            Q.set_backend(i8="auto")
            assert relerr(lay(cu(c["xq"])), c["out"]) < 5e-5          # untagged
        c = case(golden, f"lin_dorefa_w{k}_real")
        lay = _set(Q.layers.LinearDorefa(K, N, bit_width=k).cuda(), c["w"], c["b"])
        assert relerr(lay(cu(c["x"])), c["out"]) < 5e-5


def test_linear_xnor_golden(Q, golden):
    c = case(golden, "lin_xnor")
    N, K = c["w"].shape
    lay = _set(Q.layers.LinearXNOR(K, N).cuda(), c["w"], c["b"])
    with torch.no_grad():
        for mode, tol in (("fp16", 5e-4), ("bf16x2", 2e-5)):
            Q.set_xnor_mode(mode)                          # one fp16 pass (11-bit alpha*sign) vs two bf16 passes
            xq = Q.functions.QuantXnor(cu(c["x"]), 1)
            assert relerr(xq, c["xq"]) < 2e-6
            for be in ("simt", "tcgen05"):
                Q.set_backend(bf16=be)
                assert relerr(lay(xq), c["out"]) < tol, (mode, be)
                assert relerr(lay(cu(c["x"])), c["out_real"]) < 5e-5, be
        Q.set_backend(bf16="auto")
        Q.set_xnor_mode("fp16")
        # XnorNet activations feeding an integer-weight layer (fp16 exact-code route)
        lb = _set(Q.layers.LinearBin(K, N).cuda(), c["w"], c["b"])
        ref = O.linear_bin(c["xq"], c["w"], c["b"])
        assert relerr(lb(Q.functions.QuantXnor(cu(c["x"]), 1)), ref) < 2e-5
        lay.train(False)                                  # the reference raises NameError here; ours swaps
        assert relerr(lay.weight.data, O.xnor_weight(c["w"])) < 1e-6
        lay.train(True)
        assert torch.equal(lay.weight.data.cpu(), c["w"])


def test_linear_loglin_golden(Q, golden):
    with torch.no_grad():
        for dt, fsr, bw in (("lin", 7, 3), ("log", 7, 3), ("log", 2, 2)):
            c = case(golden, f"lin_loglin_{dt}_{fsr}_{bw}")
            N, K = c["w"].shape
            lay = _set(Q.layers.LinearQuant(K, N, dtype=dt, fsr=fsr, bit_width=bw).cuda(), c["w"], c["b"])
            assert relerr(lay(cu(c["x"])), c["out"]) < 5e-5


# ------------------------------------------------------------------ conv layers vs golden
CONV = {"s1p1": dict(stride=1, padding=1), "s2p0": dict(stride=2, padding=0),
        "s1p2d2": dict(stride=1, padding=2, dilation=2)}


@pytest.mark.parametrize("tag", list(CONV))
def test_conv_layers_golden(Q, golden, tag):
    kw = CONV[tag]
    Lm, F = Q.layers, Q.functions
    with torch.no_grad():
        c = case(golden, f"conv_bin_{tag}")
        lay = _set(Lm.BinConv2d(5, 7, 3, **kw).cuda(), c["w"], c["b"])
        for be in ("simt", "tcgen05"):
            Q.set_backend(i8=be)
            assert torch.equal(lay(F.BinaryConnect()(cu(c["x"]))).cpu(), c["out"]), be     # zero padding, bit exact
        Q.set_backend(i8="auto")
        assert relerr(lay(cu(c["x"])), c["out_real"]) < 5e-5
        c = case(golden, f"conv_ter_{tag}")
        lay = _set(Lm.TerConv2d(5, 7, 3, **kw).cuda(), c["w"], c["b"])
        assert torch.equal(lay(F.BinaryConnect()(cu(c["x"]))).cpu(), c["out"])
        for k in (2, 4, 8):
            c = case(golden, f"conv_dorefa_w{k}a{k}_{tag}")
            lay = _set(Lm.DorefaConv2d(5, 7, 3, bit_width=k, **kw).cuda(), c["w"], c["b"])
            assert relerr(lay(F.DorefaQuant(cu(c["x"]), k)), c["out"]) < 2e-5, k
        c = case(golden, f"conv_xnor_{tag}")
        lay = _set(Lm.XNORConv2d(5, 7, 3, **kw).cuda(), c["w"], c["b"])
        assert relerr(lay(cu(c["x"])), c["out"]) < 5e-5


def test_conv_groups(Q, golden):
    c = case(golden, "conv_bin_groups2")
    lay = _set(Q.layers.BinConv2d(6, 8, 3, padding=1, groups=2, bias=False).cuda(), c["w"])
    with torch.no_grad():
        assert torch.equal(lay(Q.functions.BinaryConnect()(cu(c["x"]))).cpu(), c["out"])


# ------------------------------------------------------------------ raw C-ABI contractions: exact accumulators
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (33, 24, 70), (128, 256, 128), (200, 300, 1000), (257, 513, 4100),
                                   (512, 1024, 1024)])
@pytest.mark.parametrize("backend", ["simt", "tcgen05"])
def test_gemm_i8_accumulators_exact(Q, M, N, K, backend):
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    for a_signed, w_signed in ((True, True), (False, True), (True, False), (False, False)):
        lo_a, hi_a = (-128, 128) if a_signed else (0, 256)
        lo_w, hi_w = (-128, 128) if w_signed else (0, 256)
        a = torch.randint(lo_a, hi_a, (M, K), generator=g)
        w = torch.randint(lo_w, hi_w, (N, K), generator=g)
        ld = ops.round_up(K, 16)
        ad = torch.zeros(M, ld, dtype=torch.int8 if a_signed else torch.uint8).cuda()
        wd = torch.zeros(N, ld, dtype=torch.int8 if w_signed else torch.uint8).cuda()
        ad[:, :K] = a.to(ad.dtype).cuda(); wd[:, :K] = w.to(wd.dtype).cuda()
        acc = torch.empty(M, N, dtype=torch.int32).cuda()
        out = torch.empty(M, N).cuda()
        epi = ops.make_epi(out, ldo=N, acc_out=acc)
        ops.gemm_i8(ad, a_signed, ld, wd, w_signed, ld, M, N, K, epi,
                    L.BACKEND_SIMT if backend == "simt" else L.BACKEND_TCGEN05)
        ref = O.int_acc(a.numpy(), w.numpy())
        assert np.array_equal(acc.cpu().numpy().astype(np.int64), ref), (a_signed, w_signed)
        assert np.array_equal(out.cpu().numpy().astype(np.int64), ref) or np.abs(ref).max() >= 2 ** 24


@pytest.mark.parametrize("M,N,K", [(1, 5, 31), (33, 24, 70), (64, 1000, 4096), (300, 200, 1025)])
def test_popcount_gemms_exact(Q, M, N, K):
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    g = torch.Generator().manual_seed(K)
    x = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) * 0.7
    _, tag = ops.quant_act(x.cuda(), L.Q_SIGN, want_y=False, codes_kind=L.CODES_I8, want_bits=True, kind="sign")
    assert np.array_equal(tag.bits.cpu().numpy().view(np.uint32)[:, :(K + 31) // 32], O.sign_bits_packed(x))
    ps = ops.pack_weight(w.cuda(), "sign")
    assert np.array_equal(ps.packed.view(torch.int32)[0].cpu().numpy().view(np.uint32)[:, :(K + 31) // 32],
                          O.sign_bits_packed(w))
    acc = torch.empty(M, N, dtype=torch.int32).cuda()
    ops.gemm_b1b1(tag.bits, tag.ld_bits, ps.packed.view(torch.int32)[0], ps.ld_packed // 4, M, N, K,
                  ops.make_epi(None, ldo=N, acc_out=acc))
    assert np.array_equal(acc.cpu().numpy(), O.int_acc(O.sign_codes(x), O.sign_codes(w)))
    pt = ops.pack_weight(w.cuda(), "ternary")
    wb = pt.packed.view(torch.int32)
    ops.gemm_b1t2(tag.bits, tag.ld_bits, wb[0], wb[1], pt.ld_packed // 4, M, N, K, ops.make_epi(None, ldo=N, acc_out=acc))
    assert np.array_equal(acc.cpu().numpy(), O.int_acc(O.sign_codes(x), O.ternary_codes(w)))


@pytest.mark.parametrize("M,N,K", [(33, 24, 70), (256, 512, 1024), (130, 1000, 4096)])
@pytest.mark.parametrize("backend", ["simt", "tcgen05"])
def test_gemm_bf16_planes(Q, M, N, K, backend):
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    g = torch.Generator().manual_seed(K + 1)
    x = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g)
    _, ta = ops.quant_act(x.cuda(), L.Q_SPLIT, want_y=False, codes_kind=L.CODES_BF16X2, kind="real")
    pw = ops.pack_real_weight(w.cuda())
    out = torch.empty(M, N).cuda()
    ops.gemm_f16(ta.codes, ta.ld, ta.codes.stride(0), pw.planes, pw.ld_planes, pw.planes.stride(0),
                  [(0, 0), (1, 0), (0, 1)], M, N, K, ops.make_epi(out, ldo=N),
                  L.BACKEND_SIMT if backend == "simt" else L.BACKEND_TCGEN05)
    ref = x.double() @ w.double().t()
    assert relerr(out, ref) < 1e-4          # 3 of the 4 hi/lo cross terms: ~2^-16 relative


# ------------------------------------------------------------------ BASELINE config sizes
def test_config1_linearbin_1024_b512(Q):
    """BASELINE configs[0]: BinaryNet LinearBin 1024x1024, batch 512.  Integer accumulators (bias-free output) are
    bit exact vs the oracle on every backend; with the bias the CPU sgemm folds b into its K-blocked accumulation, so
    the last bit depends on MKL's blocking -> 1e-6 relative."""
    torch.manual_seed(1234)
    lay = Q.layers.LinearBin(1024, 1024)
    lay.bias.data.uniform_(-1, 1)
    x = torch.randn(512, 1024)
    ref = O.binary_mlp_layer(x, lay.weight.data, lay.bias.data)
    ref_acc = O.binary_mlp_layer(x, lay.weight.data, None)
    lay = lay.cuda()
    nob = Q.layers.LinearBin(1024, 1024, bias=False).cuda()
    nob.weight.data.copy_(lay.weight.data)
    with torch.no_grad():
        for kw in (dict(i8="tcgen05"), dict(i8="simt"), dict(popcount=True)):
            Q.set_backend(**kw)
            assert torch.equal(nob(Q.functions.BinaryConnect()(x.cuda())).cpu(), ref_acc), kw
            assert relerr(lay(Q.functions.BinaryConnect()(x.cuda())), ref) < 1e-6, kw
        Q.set_backend(i8="auto", popcount=False)
        nob.eval()
        assert torch.equal(nob(Q.functions.BinaryConnect()(x.cuda())).cpu(), ref_acc)


def test_north_star_shape_properties(Q):
    """LinearBin 4096x4096, batch 8192 (north-star shape): size-independent checks -- a 64-row slice against the
    oracle, the column-checksum identity sum_n y[m,n] = x_m . sum_n w_n + sum b, and (bias-free) negation antisymmetry
    with integer-valued outputs."""
    torch.manual_seed(7)
    lay = Q.layers.LinearBin(4096, 4096)
    lay.bias.data.uniform_(-1, 1)
    w, b = lay.weight.data.clone(), lay.bias.data.clone()
    lay = lay.cuda().eval()
    x = torch.randn(8192, 4096)
    x[x == 0] = 1.0            # sign(-0) == sign(+0) == +1: exact zeros would (correctly) break the antisymmetry check
    with torch.no_grad():
        xq = Q.functions.BinaryConnect()(x.cuda())
        y = lay(xq)
        rows = torch.arange(0, 8192, 128)
        ref = O.binary_mlp_layer(x[rows], w, b)
        assert relerr(y[rows.cuda()], ref) < 1e-6
        ref_acc = O.binary_mlp_layer(x[rows], w, None)
        assert torch.equal((y[rows.cuda()] - lay.bias).round().cpu(), ref_acc)
        wsum = O.sign_codes(w).sum(0)                                   # [K]
        chk = torch.from_numpy(O.sign_codes(x) @ wsum).double() + b.double().sum()
        assert torch.allclose(y.double().sum(1).cpu(), chk, rtol=0, atol=0.5)
        lay.bias.data.zero_()                                           # without bias: exact antisymmetry
        y_pos = lay(Q.functions.BinaryConnect()(x.cuda()))
        y_neg = lay(Q.functions.BinaryConnect()((-x).cuda()))
        assert torch.equal(y_neg, -y_pos) and torch.equal(y_pos, torch.round(y_pos))


def test_config2_xnor_mlp_slice(Q):
    """BASELINE configs[1] topology (4096-4096-4096-1000, 1-bit W/A) at batch 256 vs the oracle."""
    torch.manual_seed(11)
    Lm, F = Q.layers, Q.functions
    dims = [4096, 4096, 4096, 1000]
    lays = [Lm.LinearXNOR(dims[i], dims[i + 1]) for i in range(3)]
    for l in lays:
        l.bias.data.uniform_(-1, 1)
    x = torch.randn(256, 4096)
    ref = O.xnor_mlp_forward(x, [l.weight.data for l in lays], [l.bias.data for l in lays])
    net = torch.nn.Sequential(*[m for l in lays for m in (F.nnQuantXnor(1), l)]).cuda()
    with torch.no_grad():
        y = net(x.cuda())
    assert relerr(y, ref) < REL


# ------------------------------------------------------------------ implicit-GEMM conv (TMA im2col) vs explicit gather vs oracle
@pytest.mark.parametrize("cfg", [
    dict(B=2, C=64, H=14, W=14, O=96, k=3, stride=1, padding=1, dilation=1, groups=1),
    dict(B=3, C=64, H=15, W=13, O=40, k=3, stride=2, padding=1, dilation=1, groups=1),
    dict(B=2, C=128, H=9, W=9, O=300, k=3, stride=1, padding=2, dilation=2, groups=1),
    dict(B=2, C=192, H=12, W=12, O=64, k=5, stride=1, padding=2, dilation=1, groups=1),
    dict(B=2, C=64, H=10, W=10, O=32, k=1, stride=2, padding=0, dilation=1, groups=1),
    dict(B=2, C=64, H=8, W=8, O=48, k=3, stride=1, padding=1, dilation=1, groups=2),
    dict(B=5, C=32, H=7, W=7, O=16, k=3, stride=1, padding=0, dilation=1, groups=1),
])
def test_implicit_conv_bit_exact(Q, cfg):
    c = dict(cfg)
    B, C, H, W, Oc, k = (c.pop(n) for n in ("B", "C", "H", "W", "O", "k"))
    g = torch.Generator().manual_seed(Oc * 7 + C)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Oc, C // c["groups"], k, k, generator=g) * 0.6
    ref_bin = O_conv(O_fn="conv_bin", x=O.binary_det(x), w=w, **c)
    ref_ter = O_conv(O_fn="conv_ter", x=O.binary_det(x), w=w, **c)
    xu = torch.rand(B, C, H, W, generator=g)
    ref_d4 = O_conv(O_fn="conv_dorefa", x=O.dorefa_quantize(xu, 4), w=w, bit_width=4, **c)
    ref_d8 = O_conv(O_fn="conv_dorefa", x=O.dorefa_quantize(xu, 8), w=w, bit_width=8, **c)
    with torch.no_grad():
        for implicit in (True, False):
            Q.set_implicit_conv(implicit)
            lay = Q.layers.BinConv2d(C, Oc, k, bias=False, **c).cuda(); lay.weight.data.copy_(w)
            assert torch.equal(lay(Q.functions.BinaryConnect()(x.cuda())).cpu(), ref_bin), implicit
            lay = Q.layers.TerConv2d(C, Oc, k, bias=False, **c).cuda(); lay.weight.data.copy_(w)
            assert torch.equal(lay(Q.functions.BinaryConnect()(x.cuda())).cpu(), ref_ter), implicit
            lay = Q.layers.DorefaConv2d(C, Oc, k, bias=False, bit_width=4, **c).cuda(); lay.weight.data.copy_(w)
            assert relerr(lay(Q.functions.DorefaQuant(xu.cuda(), 4)), ref_d4) < 2e-5, implicit
            lay = Q.layers.DorefaConv2d(C, Oc, k, bias=False, bit_width=8, **c).cuda(); lay.weight.data.copy_(w)
            assert relerr(lay(Q.functions.DorefaQuant(xu.cuda(), 8)), ref_d8) < 2e-5, implicit   # uint8 codes + patch sums
    Q.set_implicit_conv(True)


def O_conv(O_fn, x, w, **kw):
    return getattr(O, O_fn)(x, w, None, **kw)


# ------------------------------------------------------------------ code-only activation mode (fused inference chains)
def test_code_only_activations_bit_identical(Q):
    torch.manual_seed(5)
    F, Lm = Q.functions, Q.layers
    x = torch.randn(96, 256).cuda()
    nets = [
        torch.nn.Sequential(F.nnQuantXnor(1), Lm.LinearXNOR(256, 128), F.nnQuantXnor(1), Lm.LinearXNOR(128, 10)),
        torch.nn.Sequential(F.BinaryConnect(), Lm.LinearBin(256, 128), torch.nn.Hardtanh(), F.BinaryConnect(), Lm.LinearTer(128, 10)),
        torch.nn.Sequential(torch.nn.Hardtanh(0., 1.), F.nnDorefaQuant(4), Lm.LinearDorefa(256, 64, bit_width=4),
                            torch.nn.Hardtanh(0., 1.), F.nnDorefaQuant(8), Lm.LinearDorefa(64, 10, bit_width=8)),
    ]
    xi = torch.rand(3, 64, 9, 9).cuda()
    convnet = torch.nn.Sequential(F.nnDorefaQuant(4), Lm.DorefaConv2d(64, 32, 3, padding=1, bit_width=4),
                                  torch.nn.Hardtanh(0., 1.), F.BinaryConnect(), Lm.BinConv2d(32, 8, 3))
    with torch.no_grad():
        for net, inp in [(n, x) for n in nets] + [(convnet, xi)]:
            net = net.cuda().eval()
            ref = net(inp)
            with Q.code_only_activations():
                y = net(inp)
            assert torch.equal(y, ref)
        # the placeholder carries no storage: anything but a quantized layer fails loudly
        with Q.code_only_activations():
            ph = F.BinaryConnect()(x)
        assert ph.is_meta and tuple(ph.shape) == tuple(x.shape)
        with pytest.raises(Exception):
            (ph + x).sum().item()
        with pytest.raises(RuntimeError):
            Lm.LinearXNOR(256, 8).cuda()(ph)          # int8 sign codes cannot feed the alpha[k]-scaled XNOR weights
    # with autograd on, the quantizers keep returning real tensors
    with Q.code_only_activations():
        assert not F.BinaryConnect()(x.clone().requires_grad_(True)).is_meta


def test_eval_pack_survives_deepcopy_and_device_roundtrip(Q):
    """A layer in eval mode that is deep-copied (or moved) no longer matches its cached pack: it must re-derive the
    k-bit pack (from weight.org, or from the stored +-1 / ternary values) and give bit-identical outputs."""
    import copy
    torch.manual_seed(2)
    x = torch.randn(70, 96).cuda()
    for cls, kw in ((Q.layers.LinearBin, {}), (Q.layers.LinearTer, {}), (Q.layers.LinearDorefa, dict(bit_width=4))):
        lay = cls(96, 40, **kw).cuda()
        lay.weight.data.mul_(3)
        lay.eval()
        act = Q.functions.BinaryConnect() if cls is not Q.layers.LinearDorefa else Q.functions.nnDorefaQuant(4)
        with torch.no_grad():
            xin = x if cls is not Q.layers.LinearDorefa else x.abs().clamp(0, 1)
            ref = lay(act(xin))
            twin = copy.deepcopy(lay)
            if cls is not Q.layers.LinearDorefa:                 # quantized values are a fixed point -> k-bit pack again
                assert twin._current_pack().kind in ("sign", "ternary")
                assert torch.equal(twin(act(xin)), ref)
            else:                                                # DoReFa values re-enter as real weights (bf16 planes)
                assert relerr(twin(act(xin)), ref.cpu()) < 1e-4
            moved = lay.cpu().cuda()                             # weight.org stays behind on the old device
            assert torch.equal(moved(act(xin)), ref)


# ------------------------------------------------------------------ backward (STE) vs gradients of the live reference
@pytest.fixture(scope="module")
def golden_grads():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "quanttorch_ref_grads_v1.npz"))
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def _grad_case(gg, name):
    pre = name + "/"
    return {k[len(pre):]: v for k, v in gg.items() if k.startswith(pre)}


@pytest.mark.parametrize("name", ["lin_bin", "lin_ter", "lin_dorefa1", "lin_dorefa2", "lin_dorefa4", "lin_xnor",
                                  "conv_bin", "conv_ter", "conv_dorefa3"])
def test_backward_matches_reference(Q, golden_grads, name):
    """Forward through the low-bit kernels, backward with the reference's STE semantics: gradients w.r.t. input, weight
    and bias equal those of the live reference (golden vectors from oracle/gen_golden_grads.py) to 1e-4."""
    c = _grad_case(golden_grads, name)
    Lm, F = Q.layers, Q.functions
    w = c["w"]
    if name == "lin_bin":
        lay, act = Lm.LinearBin(w.shape[1], w.shape[0]), F.BinaryConnect()
    elif name == "lin_ter":
        lay, act = Lm.LinearTer(w.shape[1], w.shape[0]), F.BinaryConnect()
    elif name.startswith("lin_dorefa"):
        k = int(name[-1]); lay, act = Lm.LinearDorefa(w.shape[1], w.shape[0], bit_width=k), F.nnDorefaQuant(k)
    elif name == "lin_xnor":
        lay, act = Lm.LinearXNOR(w.shape[1], w.shape[0]), F.nnQuantXnor(1)
    elif name == "conv_bin":
        lay, act = Lm.BinConv2d(4, 6, 3, stride=2, padding=1), F.BinaryConnect()
    elif name == "conv_ter":
        lay, act = Lm.TerConv2d(4, 6, 3, padding=1), F.BinaryConnect()
    else:
        lay, act = Lm.DorefaConv2d(4, 6, 3, padding=1, bit_width=3), F.nnDorefaQuant(3)
    lay = lay.cuda()
    lay.weight.data.copy_(w); lay.bias.data.copy_(c["b"])
    x = c["x"].cuda().requires_grad_(True)
    y = lay(act(x))
    assert relerr(y.detach(), c["y"]) < 5e-4
    (y * c["go"].cuda()).sum().backward()
    assert relerr(x.grad, c["gx"]) < 1e-4
    assert relerr(lay.weight.grad, c["gw"]) < 1e-4
    assert relerr(lay.bias.grad, c["gb"]) < 1e-4


def test_host_pipeline_matches_direct(Q):
    from pytorch_quantize_impls_b200.pipeline import HostPipeline
    torch.manual_seed(9)
    net = torch.nn.Sequential(Q.functions.BinaryConnect(), Q.layers.LinearBin(256, 64)).cuda().eval()
    xs = [torch.randn(128, 256).pin_memory() for _ in range(5)]
    outs = [torch.empty(128, 64).pin_memory() for _ in range(5)]
    with torch.no_grad():
        HostPipeline(net, depth=2).run(xs, outs)
        torch.cuda.synchronize()
        for xi, yo in zip(xs, outs):
            assert torch.equal(net(xi.cuda()).cpu(), yo)


def test_stochastic_ops_statistics(Q):
    """BinaryConnectStochastic / TernaryConnectStochastic: device RNG, so parity is statistical, exactly as the reference's
    own tests do (BinaryNet/function_test.py:36-44, Terner/function_test.py:30-46)."""
    x = torch.tensor([0.5, -0.5, 0.0, 0.9, -0.2, 2.0, -3.0]).cuda()
    acc_b = torch.zeros_like(x); acc_t = torch.zeros_like(x)
    n = 2000
    for _ in range(n):
        acc_b += Q.functions.BinaryConnectStochastic.apply(x)
        acc_t += Q.functions.TernaryConnectStochastic.apply(x)
    exp_b = torch.clamp(x, -1, 1)                       # E[+-1] = 2 hardsigmoid(x) - 1
    exp_t = torch.sign(x) * torch.clamp(x.abs(), max=1.0)
    assert float((acc_b / n - exp_b).abs().max()) < 0.08
    assert float((acc_t / n - exp_t).abs().max()) < 0.08


def test_edge_shapes(Q):
    """Empty batch, single row, K smaller than a word, non-contiguous input."""
    lay = Q.layers.LinearBin(5, 3).cuda()
    act = Q.functions.BinaryConnect()
    with torch.no_grad():
        assert lay(act(torch.zeros(0, 5).cuda())).shape == (0, 3)
        x = torch.randn(1, 5)
        assert relerr(lay(act(x.cuda())), O.linear_bin(O.binary_det(x), lay.weight.data.cpu(), lay.bias.data.cpu())) < 1e-6
        xt = torch.randn(5, 7).cuda().t()                # non-contiguous [7, 5]
        ref = O.linear_bin(O.binary_det(xt.cpu()), lay.weight.data.cpu(), lay.bias.data.cpu())
        assert relerr(lay(act(xt)), ref) < 1e-6
        x3 = torch.randn(2, 4, 5).cuda()                 # leading dims, as F.linear accepts
        assert lay(x3).shape == (2, 4, 3)
        conv = Q.layers.BinConv2d(3, 4, 3).cuda()
        assert conv(torch.randn(0, 3, 8, 8).cuda()).shape == (0, 4, 6, 6)
