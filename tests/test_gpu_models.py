"""GPU: the BASELINE network configs against their CPU twins built on the oracle (same builder, same state_dict,
BatchNorm statistics calibrated on the twin so that activations spread over the quantizer range).

Two checks per network:
  * teacher-forced, layer by layer: every quantized layer of the GPU net is fed the input its twin saw (re-tagged
    through the GPU activation quantizer when that input is already quantized, so the integer route is the one
    exercised) and must match the twin's output to 1e-3 of max|y| (observed ~1e-6);
  * end to end: outputs agree in direction (cosine > 0.98).  Bit-level agreement end to end is not a property any two
    fp32 implementations of a deep k-bit network have: a 1e-7 difference in a pre-activation that sits on a rounding
    boundary of the next activation quantizer flips a code, and the flips cascade (the reference on its own CUDA path
    diverges from its CPU path the same way).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle_lib  # noqa: E402
import quanttorch_oracle as O  # noqa: E402

QLAYERS = ("LinearBin", "BinConv2d", "LinearTer", "TerConv2d", "LinearDorefa", "DorefaConv2d", "LinearXNOR", "XNORConv2d")


def _calibrate_bn(twin, x):
    bns = [m for m in twin.modules() if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d))]
    for m in bns:
        m.momentum = 1.0
        m.train()
    with torch.no_grad():
        twin(x)
    for m in bns:
        m.eval()
        m.bias.data.fill_(0.5); m.weight.data.fill_(0.25)     # centre the pre-activations inside the [0, 1] clamp


def _build(builder, x, **kw):
    from pytorch_quantize_impls_b200 import nets
    torch.manual_seed(3)
    twin = getattr(nets, builder)(lib=oracle_lib, **kw).eval()
    _calibrate_bn(twin, x)
    torch.manual_seed(3)
    net = getattr(nets, builder)(**kw)
    net.load_state_dict(twin.state_dict())
    return net.cuda().eval(), twin


def rel(y, ref):
    return float((y - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _check(builder, x, act_quant, **kw):
    """act_quant(t, on_gpu) -> quantized t: the network's activation quantizer (used to re-tag quantized inputs)."""
    net, twin = _build(builder, x, **kw)
    rec = {}
    hooks = []
    for name, m in twin.named_modules():
        if type(m).__name__ in ("L", "C"):                      # oracle_lib quantized Linear / Conv2d
            hooks.append(m.register_forward_hook(lambda mod, inp, out, name=name: rec.__setitem__(name, (inp[0], out))))
    with torch.no_grad():
        ref = twin(x)
        y = net(x.cuda()).cpu()
    for h in hooks:
        h.remove()
    gpu_mods = dict(net.named_modules())
    worst = 0.0
    assert len(rec) >= 3
    with torch.no_grad():
        for name, (xin, yref) in rec.items():
            lay = gpu_mods[name]
            assert type(lay).__name__ in QLAYERS
            xg = xin.cuda()
            if act_quant is not None and torch.equal(act_quant(xin, False), xin):
                xg = act_quant(xg, True)                         # already-quantized input: re-tag on the device
                assert torch.equal(xg.cpu(), xin)
            r = rel(lay(xg).cpu(), yref)
            worst = max(worst, r)
            assert r < 1e-3, (name, r)
    cos = torch.nn.functional.cosine_similarity(y.flatten(), ref.flatten(), dim=0).item()
    assert cos > 0.98, cos
    return worst, cos


def _dorefa_q(k):
    import pytorch_quantize_impls_b200 as Q
    return lambda t, gpu: Q.functions.DorefaQuant(t, k) if gpu else O.dorefa_quantize(t, k)


def test_xnor_mlp_small():
    net, twin = _build("xnor_mlp", torch.randn(64, 512), dims=(512, 384, 256, 10))
    x = torch.randn(64, 512)
    with torch.no_grad():
        assert rel(net(x.cuda()).cpu(), twin(x)) < 1e-3


def test_alexnet_dorefa_w4a4():
    worst, cos = _check("alexnet_dorefa", torch.rand(2, 3, 224, 224), _dorefa_q(4), bit_width=4)
    print("alexnet layerwise worst rel", worst, "e2e cosine", cos)


def test_resnet18_ternary_a8():
    worst, cos = _check("resnet18_ternary", torch.rand(2, 3, 224, 224), _dorefa_q(8), act_bits=8)
    print("resnet18 layerwise worst rel", worst, "e2e cosine", cos)


def test_vgg_dorefa_w8a8():
    worst, cos = _check("vgg_dorefa", torch.rand(4, 3, 32, 32), _dorefa_q(8), bit_width=8)
    print("vgg layerwise worst rel", worst, "e2e cosine", cos)


@pytest.mark.parametrize("builder,x,kw", [
    ("alexnet_dorefa", torch.rand(2, 3, 224, 224), dict(bit_width=4)),
    ("resnet18_ternary", torch.rand(2, 3, 96, 96), dict(act_bits=8)),
    ("vgg_dorefa", torch.rand(4, 3, 32, 32), dict(bit_width=8)),
])
def test_fuse_inference_matches_unfused(builder, x, kw):
    """fusion.fuse_inference (BatchNorm+clamp+quantizer in one pass) and code-only activations against the plain GPU
    graph: every fused module, fed the input the plain graph saw, reproduces the plain codes up to one level on at most
    0.1 % of the elements (fma vs three roundings); end-to-end outputs agree in direction."""
    import pytorch_quantize_impls_b200 as Q
    net, _ = _build(builder, x, **kw)
    fused, _ = _build(builder, x, **kw)          # same seeds -> same weights
    fused = Q.fuse_inference(fused)
    n_fused = [m for m in fused.modules() if isinstance(m, Q.FusedBNActQuant)]
    assert len(n_fused) >= 3
    with torch.no_grad():
        ref = net(x.cuda())
        worst = 0.0
        for fm in n_fused:        # each fused module against its own un-fused composition on a spread-out input
            C = fm.bn.num_features
            if isinstance(fm.bn, torch.nn.BatchNorm1d):
                xin = torch.rand(64, C).cuda() * 3 - 1
            else:
                xin = torch.rand(8, C, 6, 6).cuda() * 3 - 1
            xin = xin * fm.bn.running_var.sqrt().view(1, -1, *([1] * (xin.dim() - 2))) + fm.bn.running_mean.view(1, -1, *([1] * (xin.dim() - 2)))
            y_plain = fm._compose(xin)
            y_fused = fm(xin)
            n = 2 ** fm.quant._qt_spec[1] - 1 if fm.quant._qt_spec[0] == "dorefa" else 1
            d = ((y_fused - y_plain) * n).abs()
            assert float(d.max()) <= 1.0 + 1e-4
            frac = float((d > 0.5).float().mean())
            worst = max(worst, frac)
            assert frac <= 1e-3, frac
            assert y_fused._qt_codes is not None
        with Q.code_only_activations():
            y = fused(x.cuda())
    cos = torch.nn.functional.cosine_similarity(y.flatten(), ref.flatten(), dim=0).item()
    assert cos > 0.98, cos
