"""GPU: the fused inter-layer chain.  `fuse_inference` moves `[BatchNorm] -> [clamp] -> activation quantizer`
into the tcgen05 epilogue of the layer in front of it (include/qtb200.h QtRequant): the epilogue writes the NEXT layer's
low-bit operand and the fp32 activation never reaches HBM.

Checked here, through the public modules (which call the C ABI):
  * the codes the epilogue writes are bit-identical to the codes the stand-alone quantizer kernel derives from the fp32
    output of the same layer (sign / ternary / DoReFa-k, int8, uint8 and e2m1 lanes, ragged M / N, 240- and 128-wide tiles);
  * XnorNet: sign codes identical, the row mean (sum of per-tile partial sums, fixed order) within 1e-6 of the fp64-accumulated
    mean, final logits within 1e-3 of the CPU oracle (the north-star tolerance);
  * with a BatchNorm folded into the epilogue scale/bias, at most 0.1 % of the codes move, by one level (one rounding instead of
    three -- the property tests/test_gpu_models.py states for the BN+clamp+quantizer kernel);
  * conv -> BN -> clamp -> quantizer -> conv chains write channels-last codes straight from the implicit-GEMM epilogue.
"""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

import quanttorch_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def Q():
    import pytorch_quantize_impls_b200 as Q
    return Q


def _codes_as_int(tag):
    """Decode an ActCodes operand to an int32 [rows, cols] tensor."""
    from pytorch_quantize_impls_b200 import _lib as L
    c = tag.codes
    if tag.codes_kind == L.CODES_F4:
        lut = torch.tensor([0, 0, 1, 0, 2, 3, 4, 6, 0, 0, -1, 0, -2, -3, -4, -6], dtype=torch.int32, device=c.device)
        b = c.to(torch.int32)
        lo, hi = lut[b & 15], lut[(b >> 4) & 15]
        full = torch.stack([lo, hi], dim=-1).reshape(c.shape[0], -1)
        return full[:, :tag.cols]
    if c.dim() == 4:     # channels-last conv codes [B, H, W, C]
        return c.to(torch.int32)
    return c[:, :tag.cols].to(torch.float32).to(torch.int32)


def _mk(Q, name, *a, **k):
    lay = getattr(Q.layers, name)(*a, **k)
    lay.bias.data.uniform_(-1, 1)
    return lay


CASES = [
    # (input quantizer, layer ctor, following quantizer, M, K, N)
    ("sign", lambda Q, K, N: _mk(Q, "LinearBin", K, N), "sign", 300, 512, 1000),
    ("sign", lambda Q, K, N: _mk(Q, "LinearBin", K, N), "ternary", 129, 96, 384),
    ("ternary", lambda Q, K, N: _mk(Q, "LinearTer", K, N), "sign", 64, 1000, 72),
    ("dorefa4", lambda Q, K, N: _mk(Q, "LinearDorefa", K, N, bit_width=4), "dorefa4", 200, 512, 520),
    ("dorefa8", lambda Q, K, N: _mk(Q, "LinearDorefa", K, N, bit_width=8), "dorefa8", 257, 256, 136),
    ("dorefa2", lambda Q, K, N: _mk(Q, "LinearDorefa", K, N, bit_width=2), "dorefa2", 130, 320, 250),
    ("sign", lambda Q, K, N: _mk(Q, "LinearBin", K, N), "dorefa4", 100, 256, 256),
]


def _quantizer(Q, name):
    F = Q.functions
    if name == "sign":
        return F.BinaryConnect()
    if name == "ternary":
        return F.TernaryConnect()
    if name.startswith("dorefa"):
        return F.nnDorefaQuant(int(name[6:]))
    if name == "xnor":
        return F.nnQuantXnor(1)
    raise ValueError(name)


def _input(name, M, K):
    g = torch.Generator().manual_seed(11)
    if name.startswith("dorefa"):
        return torch.rand(M, K, generator=g)
    return torch.randn(M, K, generator=g)


@pytest.mark.parametrize("qin,mk,qout,M,K,N", CASES)
def test_epilogue_codes_equal_quantizer_codes(Q, qin, mk, qout, M, K, N):
    torch.manual_seed(5)
    lay = mk(Q, K, N)
    if qout.startswith("dorefa"):      # bring the pre-activations into the quantizer's [0, 1] contract
        clamp = nn.Hardtanh(0.0, 1.0)
        lay.bias.data.uniform_(0.2, 0.8)
    else:
        clamp = None
    mods = [_quantizer(Q, qin), lay] + ([clamp] if clamp is not None else []) + [_quantizer(Q, qout)]
    # a consumer layer so that fuse_inference can pick the lane format the consumer reads
    consumer = _mk(Q, "LinearDorefa", N, 64, bit_width=int(qout[6:])) if qout.startswith("dorefa") else _mk(Q, "LinearBin", N, 64)
    net = nn.Sequential(*mods, consumer).cuda().eval()
    x = _input(qin, M, K).cuda()
    with torch.no_grad():
        with Q.code_only_activations():
            h = x
            for m in list(net)[:-1]:
                h = m(h)
            ref_tag = h._qt_codes
            y_ref = net[-1](h)
        fused = Q.fuse_inference(net)
        assert isinstance(fused[1], Q.FusedLayerQuant)
        with Q.code_only_activations():
            h2 = fused[1](fused[0](x))
            assert h2.is_meta
            tag = h2._qt_codes
            y = fused[2](h2)
    assert tag.codes_kind == ref_tag.codes_kind
    assert torch.equal(_codes_as_int(tag), _codes_as_int(ref_tag))
    if tag.overflow is not None:
        assert int(tag.overflow.item()) == 0
    assert torch.equal(y, y_ref)          # the consumer sees the same operand (+ the same row sums): bit-identical output


def test_xnor_chain_matches_unfused_and_oracle(Q):
    torch.manual_seed(7)
    dims = (512, 1000, 384, 100)
    lays = [Q.layers.LinearXNOR(dims[i], dims[i + 1]) for i in range(3)]
    for l in lays:
        l.bias.data.uniform_(-1, 1)
    mods = []
    for l in lays:
        mods += [Q.functions.nnQuantXnor(1), l]
    x = torch.randn(300, dims[0])
    # CPU oracle of the same chain
    h = x
    for l in lays:
        h = O.linear_xnor(O.xnor_act(h, 1), l.weight.data, l.bias.data)
    net = nn.Sequential(*mods).cuda().eval()
    with torch.no_grad():
        with Q.code_only_activations():
            t = net[0](x.cuda())
            t = net[1](t)
            ref_mid = net[2](t)._qt_codes            # quantizer kernel on the fp32 output of layer 1
            y_plain = net(x.cuda())
        fused = Q.fuse_inference(net)
        assert [type(m).__name__ for m in fused] == ["fronteur", "FusedLayerQuant", "FusedLayerQuant", "LinearXNOR"]
        with Q.code_only_activations():
            mid = fused[1](fused[0](x.cuda()))._qt_codes
            y = fused(x.cuda())
    assert mid.row_parts > 0
    assert torch.equal(mid.codes[:, :mid.cols].float(), ref_mid.codes[:, :ref_mid.cols].float())
    mean = mid.row_scale[:mid.row_parts].sum(0) * mid.row_mul
    assert float((mean - ref_mid.row_scale).abs().max() / ref_mid.row_scale.abs().max()) < 1e-5
    rel_plain = float((y - y_plain).abs().max() / y_plain.abs().max())
    rel_oracle = float((y.cpu() - h).abs().max() / h.abs().max())
    print("xnor fused chain: vs unfused", rel_plain, "vs oracle", rel_oracle)
    assert rel_plain < 1e-4
    assert rel_oracle < 1e-3


def test_batchnorm_folded_into_epilogue(Q):
    torch.manual_seed(9)
    M, K, N = 256, 512, 512
    lay = _mk(Q, "LinearDorefa", K, N, bit_width=4)
    bn = nn.BatchNorm1d(N)
    bn.running_mean.uniform_(-0.1, 0.1); bn.running_var.uniform_(0.5, 1.5)
    bn.weight.data.uniform_(0.3, 0.6); bn.bias.data.uniform_(0.3, 0.6)
    net = nn.Sequential(Q.functions.nnDorefaQuant(4), lay, bn, nn.Hardtanh(0.0, 1.0), Q.functions.nnDorefaQuant(4),
                        _mk(Q, "LinearDorefa", N, 64, bit_width=4)).cuda().eval()
    x = torch.rand(M, K).cuda()
    with torch.no_grad():
        with Q.code_only_activations():
            h = x
            for m in list(net)[:-1]:
                h = m(h)
            ref = _codes_as_int(h._qt_codes)
        fused = Q.fuse_inference(net)
        with Q.code_only_activations():
            got = _codes_as_int(fused[1](fused[0](x))._qt_codes)
    d = (got - ref).abs()
    assert int(d.max()) <= 1
    assert float((d > 0).float().mean()) <= 1e-3
    assert ref.float().std() > 2.0          # the codes really spread over the 16 levels


def test_conv_chain_writes_channels_last_codes(Q):
    torch.manual_seed(13)
    k = 4
    c1 = Q.layers.DorefaConv2d(64, 96, kernel_size=3, padding=1, bit_width=k)
    c2 = Q.layers.DorefaConv2d(96, 64, kernel_size=3, padding=1, stride=2, bit_width=k)
    bn = nn.BatchNorm2d(96)
    bn.running_mean.uniform_(-0.1, 0.1); bn.running_var.uniform_(0.5, 1.5)
    bn.weight.data.uniform_(0.3, 0.6); bn.bias.data.uniform_(0.3, 0.6)
    net = nn.Sequential(Q.functions.nnDorefaQuant(k), c1, bn, nn.Hardtanh(0.0, 1.0), Q.functions.nnDorefaQuant(k), c2).cuda().eval()
    x = torch.rand(5, 64, 14, 14).cuda()
    with torch.no_grad():
        with Q.code_only_activations():
            h = x
            for m in list(net)[:-1]:
                h = m(h)
            ref = _codes_as_int(h._qt_codes)
            y_ref = net[-1](h)
        fused = Q.fuse_inference(net)
        assert isinstance(fused[1], Q.FusedLayerQuant)
        with Q.code_only_activations():
            t = fused[1](fused[0](x))
            assert t.is_meta and t._qt_codes.layout == "nhwc"
            got = _codes_as_int(t._qt_codes)
            y = fused[2](t)
    d = (got - ref).abs()
    assert int(d.max()) <= 1
    assert float((d > 0).float().mean()) <= 1e-3
    assert float((y - y_ref).abs().max() / y_ref.abs().max()) < 2e-2      # a handful of one-level code moves


def test_fused_module_is_plain_composition_outside_code_only_mode(Q):
    torch.manual_seed(3)
    net = nn.Sequential(Q.functions.BinaryConnect(), _mk(Q, "LinearBin", 256, 256), Q.functions.BinaryConnect(),
                        _mk(Q, "LinearBin", 256, 32)).cuda().eval()
    x = torch.randn(64, 256).cuda()
    with torch.no_grad():
        ref = net(x)
        fused = Q.fuse_inference(net)
        y = fused(x)                     # drop-in mode: the caller may look at every fp32 activation
        mid = fused[1](fused[0](x))
        assert not mid.is_meta and mid.shape == (64, 256) and set(mid.unique().tolist()) <= {-1.0, 1.0}
    assert torch.equal(y, ref)
    # autograd on: plain composition as well
    xg = x.clone().requires_grad_(True)
    net.train()
    out = fused(xg)
    out.sum().backward()
    assert xg.grad is not None


def test_one_pass_xnor_quantizer_partial_sums(Q):
    """Code-only XnorNet quantizer on long rows: ONE read of x, 1024-column chunks each leaving a partial row sum that the
    consumer epilogue adds in order; codes = torch.sign(x), mean within fp32 rounding of the fp64-accumulated mean, and the
    layer output within 1e-6 of the drop-in (two-pass) mode."""
    torch.manual_seed(21)
    x = torch.randn(200, 4096).cuda()
    x[3, :7] = 0.0
    lay = Q.layers.LinearXNOR(4096, 264).cuda().eval()
    q = Q.functions.nnQuantXnor(1)
    with torch.no_grad():
        full = q(x)
        y_ref = lay(full)
        with Q.code_only_activations():
            ph = q(x)
            tag = ph._qt_codes
            y = lay(ph)
    assert ph.is_meta and tag.row_parts == 4 and tuple(tag.row_scale.shape) == (4, 200)
    assert torch.equal(tag.codes[:, :4096].float(), torch.sign(x))
    mean = tag.row_scale.sum(0) * tag.row_mul
    ref_mean = x.double().mean(1).float()
    assert float((mean - ref_mean).abs().max()) <= 1e-6 * float(ref_mean.abs().max()) + 1e-9
    assert float((y - y_ref).abs().max() / y_ref.abs().max()) < 1e-6


@pytest.mark.parametrize("qname", ["sign", "ternary", "dorefa2", "dorefa4", "dorefa8", "xnor"])
def test_lean_code_only_quantizers_match_generic_kernel(Q, qname):
    """cols % 1024 == 0 in code-only mode takes the lean streaming kernels (act_quant_lean_kernel); their operand must be
    bit-identical to what the generic kernel emits in drop-in mode (codes, row sums, overflow flag), including edge values."""
    torch.manual_seed(31)
    M, K = 70, 2048
    x = _input(qname, M, K).cuda()
    if qname.startswith("dorefa"):
        x[0, :8] = torch.tensor([0.0, 1.0, 0.5, 1 / 6, 0.1667, 0.8333, 1.2, -0.3]).cuda()     # ties, out-of-range (unclamped)
    else:
        x[0, :8] = torch.tensor([0.0, -0.0, 0.5, -0.5, 0.4999, -0.5001, 1e-45, float("nan")]).cuda()
    q = _quantizer(Q, qname)
    with torch.no_grad():
        ref = q(x)._qt_codes                     # generic kernel (also writes the fp32 tensor)
        with Q.code_only_activations():
            got = q(x)._qt_codes                 # lean kernel
    assert got.codes_kind == ref.codes_kind
    if qname == "xnor":
        assert torch.equal(got.codes.float(), ref.codes.float())
        mean = got.row_scale.sum(0) * got.row_mul if got.row_parts else got.row_scale
        ok = torch.isfinite(ref.row_scale)
        assert float((mean[ok] - ref.row_scale[ok]).abs().max()) <= 2e-7 * float(ref.row_scale[ok].abs().max()) + 1e-9
    else:
        assert torch.equal(_codes_as_int(got), _codes_as_int(ref))
        if ref.row_sum is not None:
            assert torch.equal(got.row_sum, ref.row_sum)
        if ref.overflow is not None:
            assert int(got.overflow.item()) == int(ref.overflow.item())


@pytest.mark.parametrize("conv", [False, True])
@pytest.mark.parametrize("with_act", [False, True])
def test_batchnorm_and_clamp_folded_into_fp32_epilogue(Q, conv, with_act):
    """layer -> BatchNorm (eval) -> [Hardtanh] with no quantizer behind it: FusedLayerBN folds the affine (and the clamp) into
    the epilogue; the fp32 result matches the three-module composition to fp32 rounding."""
    torch.manual_seed(17)
    if conv:
        lay = Q.layers.TerConv2d(64, 96, kernel_size=3, padding=1, bias=True)
        bn = nn.BatchNorm2d(96)
        q, x = Q.functions.nnDorefaQuant(8), torch.rand(4, 64, 10, 10)
    else:
        lay = Q.layers.LinearDorefa(512, 264, bit_width=4)
        bn = nn.BatchNorm1d(264)
        q, x = Q.functions.nnDorefaQuant(4), torch.rand(130, 512)
    bn.running_mean.uniform_(-0.2, 0.2); bn.running_var.uniform_(0.5, 1.5)
    bn.weight.data.uniform_(0.3, 0.9); bn.bias.data.uniform_(-0.3, 0.6)
    mods = [q, lay, bn] + ([nn.Hardtanh(0.0, 1.0)] if with_act else [])
    net = nn.Sequential(*mods).cuda().eval()
    with torch.no_grad():
        ref = net(x.cuda())
        fused = Q.fuse_inference(net)
        if conv:
            assert isinstance(fused[1], Q.FusedLayerBN) and len(fused) == 2
        else:          # quantizer -> Linear head pair: one FusedActLayer around the BatchNorm-folded layer
            assert isinstance(fused[0], Q.FusedActLayer) and isinstance(fused[0].inner, Q.FusedLayerBN) and len(fused) == 1
        y = fused(x.cuda())
    assert y.shape == ref.shape
    assert float((y - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max()))
    if with_act:
        assert float(y.min()) >= 0.0 and float(y.max()) <= 1.0


def test_operand_prefetch_is_transparent(Q):
    """fusion.OperandPrefetch launches the weight expansions of all layers on a side stream at the start of the forward: same
    results bit for bit, eagerly and under CUDA-graph replay."""
    from pytorch_quantize_impls_b200.pipeline import GraphedModule
    torch.manual_seed(23)
    lays = [Q.layers.LinearXNOR(512, 384), Q.layers.LinearXNOR(384, 256), Q.layers.LinearBin(256, 64)]
    net = nn.Sequential(Q.functions.nnQuantXnor(1), lays[0], Q.functions.nnQuantXnor(1), lays[1], Q.functions.BinaryConnect(),
                        lays[2]).cuda().eval()
    x = torch.randn(300, 512).cuda()
    with torch.no_grad():
        ref = net(x)
        pf = Q.prefetch_operands(net)
        y0 = pf(x)                      # first call: nothing to prefetch yet (kinds unknown) or kinds learnt from `ref`
        y1 = pf(x)
        assert torch.equal(y0, ref) and torch.equal(y1, ref)
        assert all(l._current_pack()._prefetch is None for l in lays)
        gm = GraphedModule(lambda t: pf(t), x.clone())
        for _ in range(3):
            yg = gm()
        assert torch.equal(yg, ref)
