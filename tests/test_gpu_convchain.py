"""GPU: the channels-last conv chain of the network configs (through the C ABI) against CPU references.

  * first-layer implicit GEMM on bf16 plane pixels (qt_image_planes + qt_conv_bf16) vs the oracle's F.conv2d of the fake-quantized
    weights on the fp32 image: <= 2e-6 of max|y| (24 significant bits of the input, exact integer weights, fp32 accumulation);
  * pooling on codes / fused fp32 pool + quantizer: bit-exact vs torch.max_pool2d;
  * residual add + clamp + next quantizer in the conv epilogue, conv -> pool -> BN -> clamp -> quantizer on codes, channels-last
    flatten into a re-ordered Linear, the banded quantizer / contraction pipeline: against the plain composition of the same
    modules (bit-exact where the arithmetic is integer, <= 1 code level on <= 0.1 % of elements where a BatchNorm is folded)."""
import pytest
import torch
import torch.nn.functional as TF
from torch import nn

pytestmark = pytest.mark.gpu

import quanttorch_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def Q():
    import pytorch_quantize_impls_b200 as Q
    assert torch.cuda.is_available()
    return Q


def rel(y, ref):
    ref = ref.double().cpu()
    return float((y.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def test_image_planes_reconstruct_the_image(Q):
    from pytorch_quantize_impls_b200 import _ops as ops
    torch.manual_seed(0)
    x = (torch.randn(3, 3, 13, 17) * 4).cuda()
    for P in (1, 2, 3):
        out = ops.image_planes(x, P, 2, 3, 13 + 4, 17 + 6).float()           # [B, Hp, Wp, 16]
        assert torch.equal(out[:, :2], torch.zeros_like(out[:, :2])) and torch.equal(out[:, :, :3], torch.zeros_like(out[:, :, :3]))
        core = out[:, 2:15, 3:20]                                               # [B, H, W, 16]
        recon = sum(core[..., p * 3:(p + 1) * 3] for p in range(P)).permute(0, 3, 1, 2)
        tol = {1: 2.0 ** -8, 2: 2.0 ** -16, 3: 2.0 ** -23}[P]
        assert float((recon - x).abs().max() / x.abs().max()) <= tol
        assert torch.equal(core[..., 3 * P:], torch.zeros_like(core[..., 3 * P:]))
    # source pixels that fall outside [Hp, Wp] are dropped
    small = ops.image_planes(x, 3, 1, 1, 8, 9).float()
    assert torch.equal(small[:, 1:8, 1:9, :3], x.permute(0, 2, 3, 1)[:, :7, :8].bfloat16().float())


def test_image_windows_hold_the_filter_rows(Q):
    from pytorch_quantize_impls_b200 import _ops as ops
    torch.manual_seed(1)
    B, C, H, W = 2, 3, 11, 19
    x = (torch.randn(B, C, H, W) * 3).cuda()
    for P, kw, sw, pw, slots in ((3, 7, 2, 3, 64), (3, 3, 1, 1, 32), (2, 3, 2, 0, 32), (1, 5, 1, 2, 16), (3, 11, 4, 2, 128)):
        ph = 2
        OW = (W + 2 * pw - kw) // sw + 1
        Hp = H + 2 * ph - 1                                  # the last padded row is never read by a strided filter: dropped
        rec = ops.image_windows(x, P, kw, sw, ph, pw, Hp, OW, slots).float()            # [B, Hp, OW, slots]
        assert rec.shape == (B, Hp, OW, slots)
        assert torch.equal(rec[..., kw * P * C:], torch.zeros_like(rec[..., kw * P * C:]))
        parts = rec[..., :kw * P * C].reshape(B, Hp, OW, kw, P, C)
        recon = parts.sum(4)                                  # [B, Hp, OW, kw, C]
        xp = TF.pad(x, (pw, pw + sw, ph, ph))                 # zero padding as the conv sees it
        want = torch.stack([xp[:, :, :Hp, kx:kx + (OW - 1) * sw + 1:sw] for kx in range(kw)], -1)      # [B, C, Hp, OW, kw]
        want = want.permute(0, 2, 3, 4, 1)
        tol = {1: 2.0 ** -8, 2: 2.0 ** -16, 3: 2.0 ** -23}[P]
        assert float((recon - want).abs().max() / x.abs().max()) <= tol
        hi = parts[:, :, :, :, 0]
        assert torch.equal(hi, want.bfloat16().float())


FIRST = [  # Cin, O, k, stride, pad, dil, H, W
    (3, 64, 3, 1, 1, 1, 32, 32), (3, 64, 7, 2, 3, 1, 64, 64), (3, 192, 11, 4, 2, 1, 99, 99), (5, 32, 5, 1, 2, 1, 20, 23),
    (3, 64, 3, 3, 0, 1, 31, 29), (1, 32, 3, 1, 2, 2, 18, 18), (3, 40, 7, 2, 3, 1, 37, 41), (8, 32, 3, 2, 1, 1, 21, 21),
    (3, 64, 4, 4, 0, 1, 32, 32)]


@pytest.mark.parametrize("shape", FIRST)
@pytest.mark.parametrize("fam", ["ter", "dorefa4", "dorefa8", "bin"])
def test_first_layer_implicit_gemm(Q, shape, fam):
    Cin, Oc, k, s, p, d, H, W = shape
    torch.manual_seed(Cin * 100 + k)
    x = torch.rand(3, Cin, H, W) * 2 - 0.5
    mk = {"ter": lambda: Q.layers.TerConv2d(Cin, Oc, k, stride=s, padding=p, dilation=d),
          "bin": lambda: Q.layers.BinConv2d(Cin, Oc, k, stride=s, padding=p, dilation=d),
          "dorefa4": lambda: Q.layers.DorefaConv2d(Cin, Oc, k, stride=s, padding=p, dilation=d, bit_width=4),
          "dorefa8": lambda: Q.layers.DorefaConv2d(Cin, Oc, k, stride=s, padding=p, dilation=d, bit_width=8)}[fam]
    lay = mk()
    if fam == "ter":
        lay.weight.data.mul_(0.7 / float(lay.weight.data.abs().max()))     # all three ternary levels occur
    lay.bias.data.uniform_(-1, 1)
    w, b = lay.weight.data.clone(), lay.bias.data.clone()
    assert fam != "ter" or float(O.ternary_det(w).abs().mean()) > 0.1
    wq = {"ter": O.ternary_det, "bin": O.binary_det, "dorefa4": lambda t: O.dorefa_weight(t, 4),
          "dorefa8": lambda t: O.dorefa_weight(t, 8)}[fam](w)
    ref = TF.conv2d(x.double(), wq.double(), b.double(), s, p, d)
    lay = lay.cuda()
    from pytorch_quantize_impls_b200 import _lib
    with torch.no_grad():
        _lib.launch_count(reset=True)
        y = lay(x.cuda())
        n_launch = _lib.launch_count()
        assert y.shape == ref.shape
        # DoReFa weights: the device tanh may move a code by one level on a rounding boundary (test_gpu_parity.test_weight_quantizer)
        assert rel(y, ref) <= (5e-6 if fam in ("ter", "bin") else 2e-3)
        Q.set_first_layer_implicit(False)
        try:
            y_old = lay(x.cuda())
        finally:
            Q.set_first_layer_implicit(True)
        assert rel(y, y_old) <= 5e-5                      # the explicit two-plane gather carries 16 significant bits
        Q.set_first_layer_windows(False)                  # plane pixels + space-to-depth folds (the route of filters whose row
        try:                                              # does not fit one record)
            y_fold = lay(x.cuda())
        finally:
            Q.set_first_layer_windows(True)
        assert rel(y_fold, ref) <= (5e-6 if fam in ("ter", "bin") else 2e-3) and rel(y, y_fold) <= 5e-6
    assert n_launch <= 8                                   # pack (stats + codes), expand, image planes, implicit GEMM: no gather


def test_first_layer_eval_mode_and_channels_last_output(Q):
    from pytorch_quantize_impls_b200 import _engine as eng
    torch.manual_seed(5)
    lay = Q.layers.TerConv2d(3, 64, 7, stride=2, padding=3, bias=False)
    lay.weight.data.mul_(0.7 / float(lay.weight.data.abs().max()))
    x = torch.rand(2, 3, 50, 46)
    ref = TF.conv2d(x.double(), O.ternary_det(lay.weight.data).double(), None, 2, 3)
    lay = lay.cuda().eval()
    with torch.no_grad():
        y1 = lay(x.cuda())
        y2 = lay(x.cuda())                                 # second call: cached plane-pixel weights
        pack = lay._current_pack()
        y3 = eng.conv2d(x.cuda(), pack, None, tuple(lay.weight.shape), 2, 3, 1, 1, out_format="nhwc")
    assert rel(y1, ref) <= 2e-6 and torch.equal(y1, y2)
    assert y3.is_contiguous(memory_format=torch.channels_last) and torch.equal(y3.contiguous(), y1)


@pytest.mark.parametrize("unsigned", [False, True])
@pytest.mark.parametrize("geo", [((2, 2), (2, 2), (0, 0)), ((3, 3), (2, 2), (0, 0)), ((3, 3), (2, 2), (1, 1)), ((3, 2), (1, 2), (1, 0))])
def test_pool_codes(Q, unsigned, geo):
    from pytorch_quantize_impls_b200 import _ops as ops
    torch.manual_seed(1)
    k, s, p = geo
    B, H, W, C = 3, 13, 11, 48
    x = torch.randint(0, 256, (B, H, W, C), dtype=torch.uint8) if unsigned else torch.randint(-128, 128, (B, H, W, C), dtype=torch.int8)
    xf = x.float().permute(0, 3, 1, 2)
    ref = TF.max_pool2d(xf, k, s, p).permute(0, 2, 3, 1)
    out = ops.pool_codes(x.cuda(), k, s, p)
    assert torch.equal(out.cpu().float(), ref)
    flags = (torch.arange(C) % 3 == 1).to(torch.uint8)
    ref_min = -TF.max_pool2d(-xf, k, s, p).permute(0, 2, 3, 1)
    mixed = torch.where(flags.bool().view(1, 1, 1, C), ref_min, ref)
    out = ops.pool_codes(x.cuda(), k, s, p, use_min=flags.cuda())
    assert torch.equal(out.cpu().float(), mixed)


def test_pool_quant_f32(Q):
    from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
    torch.manual_seed(2)
    x = torch.rand(2, 64, 15, 14)
    x_cl = x.cuda().contiguous(memory_format=torch.channels_last)
    for k, s, p in (((3, 3), (2, 2), (1, 1)), ((2, 2), (2, 2), (0, 0))):
        ref = TF.max_pool2d(x, k, s, p)
        out, codes, ovf = ops.pool_quant_f32(x_cl, k, s, p, want_out=True, mode=L.Q_DOREFA, bit_width=8, codes_kind=L.CODES_U8)
        assert out.is_contiguous(memory_format=torch.channels_last) and torch.equal(out.cpu(), ref)
        cref = torch.round(255 * ref).permute(0, 2, 3, 1)
        assert torch.equal(codes.cpu().float(), cref) and int(ovf.item()) == 0
        only, none, _ = ops.pool_quant_f32(x_cl, k, s, p)
        assert none is None and torch.equal(only.cpu(), ref)


def _calibrated_bn(C, two_d=True, negative=False):
    bn = (nn.BatchNorm2d if two_d else nn.BatchNorm1d)(C)
    bn.running_mean.uniform_(-0.2, 0.2)
    bn.running_var.uniform_(0.5, 1.5)
    bn.weight.data.uniform_(0.05, 0.3)
    if negative:
        bn.weight.data[::3] *= -1                          # a third of the channels pool as MIN
    bn.bias.data.uniform_(0.3, 0.7)
    return bn.eval()


def _code_diff(tag_a, tag_b):
    a, b = tag_a.codes.float(), tag_b.codes.float()
    d = (a - b).abs()
    return float(d.max()), float((d > 0).float().mean())


@pytest.mark.parametrize("negative", [False, True])
@pytest.mark.parametrize("k,mid", [(4, 64), (8, 64), (4, 192), (8, 192)])
def test_conv_pool_bn_quant_on_codes(Q, k, mid, negative):
    """conv -> MaxPool -> BatchNorm -> Hardtanh(0, 1) -> quantizer: requant epilogue + pool on codes vs the plain modules.
    mid = 192: the codes are written with a 256-channel pitch (zero pad channels) for the next conv's 128-byte K blocks."""
    torch.manual_seed(7 + k)
    F_, L_ = Q.functions, Q.layers
    net = nn.Sequential(F_.nnDorefaQuant(k), L_.DorefaConv2d(32, mid, 3, padding=1, bit_width=k), nn.MaxPool2d(3, 2, 1),
                        _calibrated_bn(mid, negative=negative), nn.Hardtanh(0.0, 1.0), F_.nnDorefaQuant(k),
                        L_.DorefaConv2d(mid, 32, 3, padding=1, bit_width=k)).cuda().eval()
    x = torch.rand(4, 32, 17, 19).cuda()
    with torch.no_grad():
        ref_codes = net[:6](x)._qt_codes                   # plain graph: NCHW fp32 + channels-last code tag
        ref = net(x)
        fused = Q.fuse_inference(net)
        assert [type(m).__name__ for m in fused] == ["fronteur", "FusedLayerPoolQuant", "DorefaConv2d"]
        with Q.code_only_activations():
            h = fused[1](fused[0](x))
            pitch = 256 if mid == 192 else mid
            assert h.is_meta and h.shape[1] == mid and h._qt_codes.codes.shape[:3] == ref_codes.codes.shape[:3]
            assert h._qt_codes.codes.shape[3] == pitch
            got = h._qt_codes.codes
            assert torch.equal(got[..., mid:], torch.zeros_like(got[..., mid:]))         # pad channels: zero codes
            d = (got[..., :mid].float() - ref_codes.codes.float()).abs()
            mx, frac = float(d.max()), float((d > 0).float().mean())
            y = fused(x)
    assert mx <= 1 and frac <= 1e-3                        # folded BatchNorm: one rounding instead of three
    assert rel(y, ref) <= 2e-2


def test_flatten_codes_into_reordered_linear(Q):
    torch.manual_seed(11)
    F_, L_ = Q.functions, Q.layers
    for k in (4, 8):
        net = nn.Sequential(F_.nnDorefaQuant(k), L_.DorefaConv2d(32, 32, 3, padding=1, bit_width=k), nn.MaxPool2d(2),
                            nn.Hardtanh(0.0, 1.0), F_.nnDorefaQuant(k), nn.Flatten(),
                            L_.LinearDorefa(32 * 4 * 5, 24, bit_width=k)).cuda().eval()
        x = torch.rand(6, 32, 8, 10).cuda()
        with torch.no_grad():
            ref = net(x)
            fused = Q.fuse_inference(net)
            assert [type(m).__name__ for m in fused] == ["fronteur", "FusedLayerPoolQuant", "FlattenCodes", "LinearDorefa"]
            with Q.code_only_activations():
                y = fused(x)
            assert fused[3]._in_perm == (32, 4, 5)
            y2 = fused(x)                                  # drop-in mode after the re-ordering: same result through fp32 tensors
        # no BatchNorm is folded here: clamp + quantizer in the epilogue reproduce the plain codes exactly
        assert rel(y, ref) <= 2e-5 and rel(y2, ref) <= 2e-5


def test_residual_epilogue_matches_composition(Q):
    """conv2 epilogue = BatchNorm + residual add + clamp + fp32 channels-last store + next quantizer codes."""
    from pytorch_quantize_impls_b200 import _engine as eng, _lib as L, fusion
    torch.manual_seed(13)
    F_ = Q.functions
    conv = Q.layers.TerConv2d(64, 64, 3, padding=1, bias=False)
    conv.weight.data.mul_(12.0)
    conv = conv.cuda().eval()
    bn = _calibrated_bn(64).cuda()
    x = torch.rand(3, 64, 12, 10).cuda()
    res = torch.rand(3, 64, 12, 10).cuda()
    with torch.no_grad():
        xq = F_.nnDorefaQuant(8)(x)
        ref = torch.clamp(bn(conv(xq)) + res, 0.0, 1.0)
        mul, add = fusion._bn_affine(bn)
        spec = eng.RequantSpec(L.Q_DOREFA, "dorefa", bit_width=8, lo=0.0, hi=1.0, col_mul=mul, col_add=add)
        spec.force_8bit = True
        res_cl = res.contiguous(memory_format=torch.channels_last)
        with Q.code_only_activations():
            y = conv._forward_requant(F_.nnDorefaQuant(8)(x), spec, out_format="nhwc", residual=res_cl, keep_out=True)
        assert y.is_contiguous(memory_format=torch.channels_last) and not y.is_meta
        assert float((y - ref).abs().max()) <= 2e-6
        codes = y._qt_codes.codes.float().permute(0, 3, 1, 2)
        d = (codes - torch.round(255 * ref)).abs()
        assert float(d.max()) <= 1 and float((d > 0).float().mean()) <= 1e-3
        assert torch.equal(codes, torch.round(255 * y))   # the codes are the quantizer of the value that was stored
        plain = eng.RequantSpec(-1, None, lo=0.0, hi=1.0, col_mul=mul, col_add=add)
        with Q.code_only_activations():
            y2 = conv._forward_affine(F_.nnDorefaQuant(8)(x), plain, out_format="nhwc", residual=res_cl)
        assert torch.equal(y2, y)


def test_fused_basic_blocks_match_plain_blocks(Q):
    from pytorch_quantize_impls_b200 import nets
    torch.manual_seed(17)
    lib = nets._default_lib()
    blocks = nn.Sequential(nets.TerBasicBlock(lib, 64, 64, 1, 8), nets.TerBasicBlock(lib, 64, 128, 2, 8),
                           nets.TerBasicBlock(lib, 128, 128, 1, 8))
    for m in blocks.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.uniform_(-0.5, 0.5); m.running_var.uniform_(20, 60)
            m.weight.data.uniform_(0.3, 0.6); m.bias.data.uniform_(0.2, 0.5)
        if type(m).__name__ == "TerConv2d":
            m.weight.data.mul_(10.0)
    blocks = blocks.cuda().eval()
    x = torch.rand(2, 64, 14, 14).cuda()
    with torch.no_grad():
        ref = blocks(x)
        fused = Q.fuse_inference(blocks)
        assert all(type(m).__name__ == "FusedBasicBlock" for m in fused)
        with Q.code_only_activations():
            y = fused(x)
        assert y.is_contiguous(memory_format=torch.channels_last)
        # each block: a flipped 8-bit code (folded BatchNorm rounding) moves an output by ~1/255 of a weight; direction must agree
        assert float((y - ref).abs().mean()) <= 2e-3
        cos = TF.cosine_similarity(y.flatten(), ref.flatten(), dim=0).item()
        assert cos > 0.999, cos
        # teacher-forced single block: same input -> same output up to the fold rounding
        b0 = fused[0]
        # the block in front of a down-sampling block (conv shortcut) writes codes only: its fp32 output has no reader
        assert b0._next_reads_fp32 is False and fused[1]._next_reads_fp32 is True and fused[2]._next_reads_fp32 is True
        with Q.code_only_activations():
            c0 = b0(x)
            assert c0.is_meta and c0._qt_codes is not None
            b0._next_reads_fp32 = True
            try:
                y0 = b0(x)
            finally:
                b0._next_reads_fp32 = False
        assert torch.equal(c0._qt_codes.codes, y0._qt_codes.codes)
        r0 = b0.block(x)
        assert float((y0 - r0).abs().max()) <= 5e-2 and float(((y0 - r0).abs() > 1e-5).float().mean()) <= 5e-3
        assert y0._qt_codes is not None and torch.equal(y0._qt_codes.codes.float().permute(0, 3, 1, 2), torch.round(255 * y0))


@pytest.mark.parametrize("fam", ["bin", "ter", "dorefa4", "xnor"])
def test_banded_head_pair_equals_plain(Q, fam):
    """FusedActLayer (banded two-stream quantizer / contraction pipeline) vs quantizer -> layer: bit-exact."""
    torch.manual_seed(19)
    F_, L_ = Q.functions, Q.layers
    K, N, M = 1024, 520, 4096 + 300
    q, lay = {"bin": (F_.BinaryConnect(), L_.LinearBin(K, N)), "ter": (F_.TernaryConnect(), L_.LinearTer(K, N)),
              "dorefa4": (F_.nnDorefaQuant(4), L_.LinearDorefa(K, N, bit_width=4)),
              "xnor": (F_.nnQuantXnor(1), L_.LinearXNOR(K, N))}[fam]
    if fam == "ter":
        lay.weight.data.mul_(20.0)
    lay.bias.data.uniform_(-1, 1)
    net = nn.Sequential(q, lay).cuda().eval()
    x = (torch.rand(M, K) if fam == "dorefa4" else torch.randn(M, K)).cuda()
    with torch.no_grad():
        ref = net(x)
        fused = Q.fuse_inference(net)
        assert type(fused[0]).__name__ == "FusedActLayer"
        Q.set_banded_head(True)
        with Q.code_only_activations():
            y = fused(x)
            g = torch.cuda.CUDAGraph()                      # the two-stream fork / join must be capturable
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fused(x)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g):
                yg = fused(x)
            g.replay()
            torch.cuda.synchronize()
        y_small = fused(x[:100])                            # below MIN_ROWS: plain path
        Q.set_banded_head(False)
    if fam == "xnor":
        assert rel(y, ref) <= 2e-6 and rel(yg, ref) <= 2e-6   # partial row sums are added in a different order
    else:
        assert torch.equal(y, ref) and torch.equal(yg, ref)
    assert torch.equal(y_small, ref[:100]) or rel(y_small, ref[:100]) <= 2e-6


@pytest.mark.parametrize("case", [("ter", 8, 3, 1, 1), ("dorefa8", 8, 3, 1, 1), ("dorefa4", 4, 5, 2, 1), ("ter", 8, 3, 1, 2), ("dorefa8", 8, 1, 0, 1)])
def test_w_folded_conv_equals_plain_implicit_gemm(Q, case):
    """64-channel stride-1 convs read two adjacent pixels as one 128-byte pixel (two parity launches): codes, fp32 side output and
    residual add must be bit-identical to the unfolded implicit GEMM (integer accumulators, same epilogue arithmetic)."""
    from pytorch_quantize_impls_b200 import _engine as eng, _lib as L, fusion
    fam, abits, k, pad, sh = case
    torch.manual_seed(23 + k)
    F_ = Q.functions
    if fam == "ter":
        conv = Q.layers.TerConv2d(64, 64, k, stride=(sh, 1), padding=pad, bias=False)
        conv.weight.data.mul_(0.7 / float(conv.weight.data.abs().max()))
    else:
        conv = Q.layers.DorefaConv2d(64, 96, k, stride=(sh, 1), padding=pad, bit_width=int(fam[6:]))
    conv = conv.cuda().eval()
    O = conv.out_channels
    bn = _calibrated_bn(O).cuda()
    x = torch.rand(8, 64, 72, 76).cuda()
    OH = (72 + 2 * pad - k) // sh + 1
    OW = 76 + 2 * pad - k + 1
    res = torch.rand(8, O, OH, OW).cuda().contiguous(memory_format=torch.channels_last)
    mul, add = fusion._bn_affine(bn)
    outs = {}
    with torch.no_grad():
        for flag in (False, True):
            Q.set_wfold(flag)
            try:
                spec = eng.RequantSpec(L.Q_DOREFA, "dorefa", bit_width=abits, lo=0.0, hi=1.0, col_mul=mul, col_add=add)
                spec.force_8bit = True
                with Q.code_only_activations():
                    xq = F_.nnDorefaQuant(abits)(x)
                    y1 = conv._forward_requant(xq, spec)                                              # codes only
                    y2 = conv._forward_requant(F_.nnDorefaQuant(abits)(x), spec, out_format="nhwc", residual=res, keep_out=True)
                    plain = eng.RequantSpec(-1, None, lo=0.0, hi=1.0, col_mul=mul, col_add=add)
                    y3 = conv._forward_affine(F_.nnDorefaQuant(abits)(x), plain, out_format="nhwc")
                outs[flag] = (y1._qt_codes.codes.clone(), y2.clone(), y2._qt_codes.codes.clone(), y3.clone())
            finally:
                Q.set_wfold(True)
    for a, b in zip(outs[False], outs[True]):
        assert a.shape == b.shape and torch.equal(a, b)
    assert outs[True][1].is_contiguous(memory_format=torch.channels_last)


def test_fp32_head_is_batch_invariant(Q):
    """Global average pool + nn.Linear as one kernel: equals torch to fp32 accuracy, and a sample's logits do not depend on the
    batch it is part of (the property the sharded-run == single-GPU-run check of bench.py --config needs)."""
    from pytorch_quantize_impls_b200 import _ops as ops
    torch.manual_seed(11)
    lin = torch.nn.Linear(512, 10).cuda()
    x = torch.rand(96, 512, 7, 7).cuda().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y = ops.head_f32(x, lin.weight, lin.bias)
        ref = torch.nn.functional.linear(x.double().mean((2, 3)), lin.weight.double(), lin.bias.double())
        assert rel(y, ref) <= 2e-6
        for n in (1, 5, 32):
            assert torch.equal(ops.head_f32(x[:n].contiguous(memory_format=torch.channels_last), lin.weight, lin.bias), y[:n])
        y2 = ops.head_f32(x[:, :, :1, :1].reshape(96, 512), lin.weight, None)          # plain Linear, no bias
        assert rel(y2, torch.nn.functional.linear(x[:, :, 0, 0].double(), lin.weight.double())) <= 2e-6
    net = Q.fuse_inference(nets_resnet().cuda().eval())
    assert net.__dict__.get("_fused_head") is not None
    xi = torch.rand(6, 3, 64, 64).cuda()
    with torch.no_grad():
        a = net(xi)
        b = torch.cat([net(xi[:2]), net(xi[2:])])
    assert torch.equal(a, b)


def nets_resnet():
    from pytorch_quantize_impls_b200 import nets
    torch.manual_seed(3)
    return nets.resnet18_ternary()


@pytest.mark.parametrize("fam", ["bin", "ter"])
@pytest.mark.parametrize("split", [0.0, 0.5])
def test_quantizer_beside_contraction_equals_plain(Q, fam, split):
    """set_overlap_head(True): the quantizer runs on a side stream while the contraction's TMA producer follows its progress
    counters (QtActQuant.ready -> QtEpilogue.a_ready) -- same bits as one kernel after the other."""
    from pytorch_quantize_impls_b200 import _engine as eng
    torch.manual_seed(21)
    M, K, N = 4096, 2048, 384
    lay = (Q.layers.LinearBin(K, N) if fam == "bin" else Q.layers.LinearTer(K, N)).cuda().eval()
    act = Q.functions.BinaryConnect() if fam == "bin" else Q.functions.TernaryConnect()
    pair = Q.fuse_inference(torch.nn.Sequential(act, lay))
    x = torch.randn(M, K).cuda()
    old = eng.OVERLAP_SPLIT[0]
    with torch.no_grad(), Q.code_only_activations():
        ref = pair(x).clone()
        Q.set_overlap_head(True)
        eng.OVERLAP_SPLIT[0] = split
        try:
            y = pair(x)
            torch.cuda.synchronize()
        finally:
            Q.set_overlap_head(False)
            eng.OVERLAP_SPLIT[0] = old
    assert torch.equal(y, ref)
