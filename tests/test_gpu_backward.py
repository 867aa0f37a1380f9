"""GPU: the backward (STE) helpers of SURVEY.md 8f-2 -- transpose+split, clip-mask STE, and the gradient contractions of the
dense layers on the bf16 tensor-core route -- against fp64 / fp32 torch references (through the C ABI)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    import pytorch_quantize_impls_b200 as Q
    return Q


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("R,C,planes", [(70, 130, 2), (64, 64, 1), (257, 33, 3), (1, 5, 2)])
def test_transpose_split(Q, R, C, planes):
    from pytorch_quantize_impls_b200 import _ops as ops
    torch.manual_seed(R + C)
    x = (torch.randn(R, C) * 3).cuda()
    out, ld = ops.transpose_split(x, planes)
    assert out.shape == (planes, C, ld) and ld % 8 == 0 and ld >= R
    recon = out.float().sum(0)
    assert torch.equal(recon[:, R:], torch.zeros(C, ld - R, device="cuda"))          # zero padding
    tol = {1: 2.0 ** -8, 2: 2.0 ** -16, 3: 2.0 ** -23}[planes]
    assert float((recon[:, :R] - x.t()).abs().max() / x.abs().max()) <= tol
    assert torch.equal(out[0, :, :R].float(), x.t().bfloat16().float())              # hi plane = round-to-nearest bf16


def test_ste_clip(Q):
    from pytorch_quantize_impls_b200 import _ops as ops
    x = torch.tensor([0.0, 1.0, -1.0, 1.001, -1.001, 1.0011, -1.5, 0.3, float("nan")] * 5 + [2.0, 0.1, 0.2]).cuda()
    g = torch.randn_like(x)
    ref = g.clone()
    ref[torch.abs(x) > 1.001] = 0
    assert torch.equal(ops.ste_clip(g, x), ref)
    big = torch.randn(1000, 333).cuda() * 1.2
    gb = torch.randn_like(big)
    refb = gb.clone(); refb[torch.abs(big) > 1.001] = 0
    assert torch.equal(ops.ste_clip(gb, big), refb)


@pytest.mark.parametrize("M,N,K", [(300, 520, 264), (1024, 1024, 512), (5, 7, 3)])
def test_gradient_contractions_match_fp64(Q, M, N, K):
    from pytorch_quantize_impls_b200 import _engine as eng
    torch.manual_seed(M)
    g = torch.randn(M, N).cuda()
    wq = (torch.randn(N, K).sign() * torch.rand(1, K)).cuda()
    x = torch.randn(M, K).cuda()
    gi = eng.grad_input_linear(g, wq)
    gw = eng.grad_weight_linear(g, x)
    assert rel(gi, g.double() @ wq.double()) < 3e-5
    assert rel(gw, g.double().t() @ x.double()) < 3e-5


@pytest.mark.parametrize("cls,kw,act", [("LinearBin", {}, "sign"), ("LinearTer", {}, "sign"), ("LinearDorefa", dict(bit_width=4), "dorefa4"),
                                        ("LinearXNOR", {}, "xnor")])
def test_training_step_grads_match_torch_backend(Q, cls, kw, act):
    """fwd on the low-bit kernels + bwd on the bf16 tensor-core route vs the same step with fp32 torch.matmul gradients."""
    F = Q.functions
    torch.manual_seed(3)
    lay = getattr(Q.layers, cls)(384, 200, **kw).cuda()
    if cls == "LinearTer":
        lay.weight.data.mul_(12.0)          # N(0, 1/sqrt(in)) weights would all ternarise to 0 (|w| < 0.5)
    q = {"sign": F.BinaryConnect(), "dorefa4": F.nnDorefaQuant(4), "xnor": F.nnQuantXnor(1)}[act]
    x0 = (torch.rand(160, 384) if act == "dorefa4" else torch.randn(160, 384) * 0.7).cuda()
    go = torch.randn(160, 200).cuda()
    grads = {}
    for backend in ("tcgen05", "torch"):
        Q.set_grad_backend(backend)
        lay.zero_grad()
        x = x0.clone().requires_grad_(True)
        y = lay(q(x))
        (y * go).sum().backward()
        grads[backend] = (x.grad.clone(), lay.weight.grad.clone(), lay.bias.grad.clone())
    Q.set_grad_backend("tcgen05")
    for a, b in zip(grads["tcgen05"], grads["torch"]):
        assert rel(a, b) < 1e-4


@pytest.mark.parametrize("geo", [dict(stride=1, padding=1), dict(stride=2, padding=1), dict(stride=1, padding=2, dilation=2),
                                 dict(stride=1, padding=1, groups=2)])
def test_conv_gradient_contractions_match_fp64(Q, geo):
    """engine.grad_input_conv2d / grad_weight_conv2d (im2col view + bf16 hi/lo planes on tcgen05) vs fp64 autograd of F.conv2d."""
    from pytorch_quantize_impls_b200 import _engine as eng
    torch.manual_seed(11)
    groups = geo.get("groups", 1)
    x = torch.randn(6, 16, 13, 11)
    w = torch.randn(24, 16 // groups, 3, 3).sign() * torch.rand(24, 1, 1, 1)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    y = torch.nn.functional.conv2d(xd, wd, None, **geo)
    go = torch.randn(y.shape)
    y.backward(go.double())
    gi = eng.grad_input_conv2d(x.shape, w.cuda(), go.cuda(), **geo)
    gw = eng.grad_weight_conv2d(x.cuda(), w.shape, go.cuda(), **geo)
    assert gi.shape == x.shape and gw.shape == w.shape
    assert rel(gi.cpu(), xd.grad) < 3e-5 and rel(gw.cpu(), wd.grad) < 3e-5
    Q.set_grad_backend("torch")
    try:
        gi_t = eng.grad_input_conv2d(x.shape, w.cuda(), go.cuda(), **geo)
    finally:
        Q.set_grad_backend("tcgen05")
    assert rel(gi.cpu(), gi_t.cpu()) < 2e-3          # cuDNN's default conv math is TF32
