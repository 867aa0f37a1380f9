"""GPU parity of the fp4 (e2m1) low-bit route: activation / weight code emission bit-exact against the oracle's packer,
integer accumulators of the tcgen05 kind::mxf4 contraction (unit block scales) bit-exact against the int64 oracle for
sign / ternary / DoReFa-2 operands, ragged shapes, every tile width, and K large enough that a reduced-precision
accumulator would show.  All calls go through the C ABI (ctypes)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import quanttorch_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def Q():
    import pytorch_quantize_impls_b200 as Q
    assert torch.cuda.is_available()
    return Q


def _act(ops, L, x, kind):
    if kind == "sign":
        _, t = ops.quant_act(x.cuda(), L.Q_SIGN, want_y=False, codes_kind=L.CODES_F4, kind="sign")
        return t, O.sign_codes(x)
    if kind == "ternary":
        _, t = ops.quant_act(x.cuda(), L.Q_TERNARY, want_y=False, codes_kind=L.CODES_F4, kind="ternary")
        return t, O.ternary_codes(x)
    _, t = ops.quant_act(x.cuda(), L.Q_DOREFA, bit_width=2, want_y=False, codes_kind=L.CODES_F4, want_row_sum=True, kind="dorefa")
    return t, O.dorefa_act_codes(x, 2)


def _wcodes(w, kind):
    if kind == "sign":
        return O.sign_codes(w)
    if kind == "ternary":
        return O.ternary_codes(w)
    return 2 * O.dorefa_weight_codes(w, 2) - 3        # centred codes of W_q = (2c - 3)/3


@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 31), (5, 70), (64, 256), (33, 1000), (9, 4100), (16, 8192)])
def test_fp4_activation_codes_bit_exact(Q, rows, cols):
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    g = torch.Generator().manual_seed(rows * 131 + cols)
    x = torch.randn(rows, cols, generator=g)
    x[0, 0] = -0.0
    for kind in ("sign", "ternary", "dorefa"):
        xi = torch.rand(rows, cols, generator=g) if kind == "dorefa" else x
        t, codes = _act(ops, L, xi, kind)
        assert t.ld % 32 == 0 and t.codes.shape == (rows, t.ld // 2)
        assert np.array_equal(t.codes.cpu().numpy(), O.e2m1_pack(codes, t.ld)), kind
        if kind == "dorefa":
            assert np.array_equal(t.row_sum.cpu().numpy().astype(np.int64), codes.sum(1))
            assert int(t.overflow.item()) == 0
    # out-of-lane DoReFa-2 codes (|c| > 4) saturate and raise the sticky flag
    bad = torch.full((2, 8), 3.0)
    _, t = ops.quant_act(bad.cuda(), L.Q_DOREFA, bit_width=2, want_y=False, codes_kind=L.CODES_F4, kind="dorefa")
    assert int(t.overflow.item()) == 1


@pytest.mark.parametrize("n,k", [(1, 1), (7, 33), (24, 70), (300, 1000), (513, 4100)])
def test_fp4_weight_expansion_bit_exact(Q, n, k):
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    g = torch.Generator().manual_seed(n + 3 * k)
    w = torch.randn(n, k, generator=g) * 0.7
    for kind, bw in (("sign", 1), ("ternary", 1), ("dorefa", 1), ("dorefa", 2)):
        p = ops.pack_weight(w.cuda(), kind, bit_width=bw)
        e, ld = ops.expand_weight(p, L.CODES_F4)
        ref = O.sign_codes(w) if (kind, bw) == ("dorefa", 1) else _wcodes(w, kind)
        assert ld % 32 == 0 and np.array_equal(e.cpu().numpy(), O.e2m1_pack(ref, ld)), (kind, bw)


SHAPES = [(1, 1, 1), (33, 24, 70), (128, 256, 128), (200, 300, 1000), (257, 513, 4100), (512, 1024, 1024), (130, 250, 20000)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tile_n", [0, 64, 128, 240])
def test_gemm_f4_accumulators_exact(Q, M, N, K, tile_n):
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    L.set_option("f4_tile_n", tile_n)
    try:
        for akind, wkind in (("sign", "sign"), ("sign", "ternary"), ("ternary", "ternary"), ("dorefa", "dorefa"), ("ternary", "dorefa")):
            x = torch.rand(M, K, generator=g) if akind == "dorefa" else torch.randn(M, K, generator=g)
            w = torch.randn(N, K, generator=g) * 0.7
            t, ca = _act(ops, L, x, akind)
            p = ops.pack_weight(w.cuda(), wkind, bit_width=2)
            e, ldw = ops.expand_weight(p, L.CODES_F4)
            acc = torch.empty(M, N, dtype=torch.int32).cuda()
            out = torch.empty(M, N).cuda()
            bias = torch.randn(N, generator=g).cuda()
            ops.gemm_f4(t.codes, t.ld, e, ldw, M, N, K, ops.make_epi(out, ldo=N, acc_out=acc, bias=bias))
            ref = O.int_acc(ca, _wcodes(w, wkind))
            assert np.array_equal(acc.cpu().numpy().astype(np.int64), ref), (akind, wkind)
            # fp32 output = float(acc) + bias: one rounding
            assert torch.equal(out.cpu(), torch.from_numpy(ref).float() + bias.cpu()), (akind, wkind)
    finally:
        L.set_option("f4_tile_n", 0)


def test_gemm_f4_unaligned_output_and_nchw(Q):
    """Row-major output with a pitch that rules out the TMA store (ldo % 4 != 0) and the NCHW epilogue."""
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    g = torch.Generator().manual_seed(5)
    M, N, K = 192, 250, 384
    x = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g)
    t, ca = _act(ops, L, x, "sign")
    e, ldw = ops.expand_weight(ops.pack_weight(w.cuda(), "sign"), L.CODES_F4)
    ref = torch.from_numpy(O.int_acc(ca, O.sign_codes(w))).float()
    out = torch.zeros(M, N + 3).cuda()
    ops.gemm_f4(t.codes, t.ld, e, ldw, M, N, K, ops.make_epi(out, ldo=N + 3))
    assert torch.equal(out[:, :N].cpu(), ref) and float(out[:, N:].abs().sum()) == 0.0
    P = 64                                         # 3 "images" of 64 pixels
    o4 = torch.zeros(M // P, N, P).cuda()
    ops.gemm_f4(t.codes, t.ld, e, ldw, M, N, K, ops.make_epi(o4, ldo=N, out_mode=1, nchw_inner=P))
    assert torch.equal(o4.cpu(), ref.reshape(M // P, P, N).permute(0, 2, 1))


def test_gemm_f4_long_k_full_precision_accumulation(Q):
    """K = 2^20 with operand patterns whose partial sums are odd multiples (63 per 64-element MMA step, 9-valued DoReFa-2
    products): any accumulator narrower than fp32's 24 bits would drop low bits somewhere along the way."""
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    M, N, K = 16, 24, 1 << 20
    ca = np.ones((M, K), np.int8)
    ca[1::2] = -1
    ca[:, ::64] = 0                                 # 63 non-zero products per MMA K-step
    cw = np.ones((N, K), np.int8)
    cw[2::3] = -1
    cw[5] = 3; ca[7] = 3                            # 9 K (1 - 1/64) = 9.29e6 < 2^24
    cw[:, 1::4096] *= -1
    a = torch.from_numpy(O.e2m1_pack(ca, K)).cuda()
    w = torch.from_numpy(O.e2m1_pack(cw, K)).cuda()
    acc = torch.empty(M, N, dtype=torch.int32).cuda()
    ops.gemm_f4(a, K, w, K, M, N, K, ops.make_epi(None, ldo=N, acc_out=acc))
    ref = (torch.from_numpy(ca).double() @ torch.from_numpy(cw).double().t()).numpy().astype(np.int64)
    assert np.abs(ref).max() < 2 ** 24
    assert np.array_equal(acc.cpu().numpy().astype(np.int64), ref)


def test_layers_route_through_fp4_and_match_int8(Q):
    """BinaryConnect / TernaryConnect / DorefaQuant(2) -> LinearBin / LinearTer / LinearDorefa(<=2): the fp4 route (default)
    is bit-identical to the int8 route and to the oracle; a DoReFa-4 consumer re-emits int8 codes."""
    from pytorch_quantize_impls_b200 import _engine as eng, _lib as L
    torch.manual_seed(11)
    B, K, N = 200, 520, 250
    x = torch.randn(B, K)
    F, Lm = Q.functions, Q.layers
    with torch.no_grad():
        for act_name, act, lay, ref_fn in (
                ("bin", F.BinaryConnect(), Lm.LinearBin(K, N), lambda x, l: O.linear_bin(O.binary_det(x), l.weight, l.bias)),
                ("bin-ter", F.BinaryConnect(), Lm.LinearTer(K, N), lambda x, l: O.linear_ter(O.binary_det(x), l.weight, l.bias)),
                ("ter-ter", F.TernaryConnect(), Lm.LinearTer(K, N), lambda x, l: O.linear_ter(O.ternary_det(x), l.weight, l.bias))):
            lay.bias.data.uniform_(-1, 1)
            ref = ref_fn(x, lay)
            lay = lay.cuda()
            xq = act(x.cuda())
            assert eng.get_tag(xq).codes_kind == L.CODES_F4
            y4 = lay(xq)
            eng.set_fp4(False)
            try:
                xq8 = act(x.cuda())
                assert eng.get_tag(xq8).codes_kind == L.CODES_I8
                y8 = lay(xq8)
            finally:
                eng.set_fp4(True)
            assert torch.equal(y4, y8), act_name
            assert float((y4.cpu() - ref).abs().max() / ref.abs().max()) < 1e-6, act_name
        xu = torch.rand(B, K)
        for kw in (1, 2, 4, 8):
            lay = Lm.LinearDorefa(K, N, bit_width=kw)
            ref = O.linear_dorefa(O.dorefa_quantize(xu, 2), lay.weight.data, lay.bias.data, kw)
            xq = F.DorefaQuant(xu.cuda(), 2)
            assert eng.get_tag(xq).codes_kind == L.CODES_F4
            y = lay.cuda()(xq)
            assert float((y.cpu() - ref).abs().max() / ref.abs().max()) < 2e-5, kw
        # code-only chain on fp4 operands
        net = torch.nn.Sequential(F.BinaryConnect(), Lm.LinearBin(K, N), F.BinaryConnect(), Lm.LinearTer(N, 64)).cuda().eval()
        y_full = net(x.cuda())
        with eng.code_only_activations():
            y_code = net(x.cuda())
        assert torch.equal(y_full, y_code)
