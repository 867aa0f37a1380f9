"""GPU parity of the *functional* dense / conv ops through the C ABI, against golden vectors
of the live reference (oracle/gen_golden_functional.py): forward outputs and the gradients of the hand-written backward.

    a5   BinaryDense                     binary_connect.py:86-112
    a14  TernaryDense / TernaryConv2d    terner_connect.py:78-153   (torch.sign form: +-0.5 -> +-0.5, 0 -> 0)
    a20  QuantDense / QuantConv2d        dorefa_connect.py:116-199

Tolerances: +-1 activations x +-1 weights are bit-exact; real-valued activations ride the bf16 3-plane split (<= 5e-5 of
max|y|, 3 planes for image-like conv inputs); gradients <= 1e-4 (dense and conv: tcgen05 bf16 hi/lo planes)."""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def Q():
    import pytorch_quantize_impls_b200 as Q
    assert torch.cuda.is_available()
    return Q


@pytest.fixture(scope="module")
def fgold():
    z = np.load(os.path.join(ROOT, "tests", "golden", "quanttorch_ref_functional_v1.npz"))
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def fcase(g, name):
    pre = name + "/"
    return {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}


def rel(y, ref):
    ref = ref.double()
    return float((y.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def run(op, c):
    x = c["x"].cuda().requires_grad_(True)
    w = c["w"].cuda().requires_grad_(True)
    b = c["b"].cuda().requires_grad_(True) if "b" in c else None
    y = op.apply(x, w, b) if b is not None else op.apply(x, w)
    y.backward(c["go"].cuda())
    return y.detach(), x.grad, w.grad, (None if b is None else b.grad)


def check(c, res, fwd_tol, grad_tol=1e-4):
    y, gx, gw, gb = res
    if fwd_tol == 0:
        assert torch.equal(y.cpu(), c["out"])
    else:
        assert rel(y, c["out"]) <= fwd_tol
    assert rel(gx, c["gx"]) <= grad_tol and rel(gw, c["gw"]) <= grad_tol
    if "gb" in c:
        assert rel(gb, c["gb"]) <= 1e-5


CONV_KW = {"s1p1": dict(stride=1, padding=1), "s2p0": dict(stride=2, padding=0), "nobias": dict(stride=1, padding=1)}


@pytest.mark.parametrize("name,tol", [("binary_dense", 5e-5), ("binary_dense_nobias", 5e-5), ("binary_dense_quant_in", 0)])
def test_binary_dense(Q, fgold, name, tol):
    c = fcase(fgold, name)
    check(c, run(Q.functions.BinaryDense, c), tol)


def test_binary_dense_on_tagged_sign_codes_is_bit_exact(Q, fgold):
    """BinaryConnect() upstream: the op contracts on the 1-bit operand; integer accumulators + bias = the reference bit for bit."""
    c = fcase(fgold, "binary_dense")
    xq = Q.functions.BinaryConnectDeterministic.apply(c["x"].cuda())
    y = Q.functions.BinaryDense.apply(xq, c["w"].cuda(), c["b"].cuda())
    c2 = fcase(fgold, "binary_dense_quant_in")       # same x, w, b with the reference's BinaryConnect in front
    assert torch.equal(xq.cpu(), c2["x"]) and torch.equal(y.cpu(), c2["out"])


@pytest.mark.parametrize("name", ["ternary_dense", "ternary_dense_nobias"])
def test_ternary_dense_keeps_half_ties(Q, fgold, name):
    c = fcase(fgold, name)
    res = run(Q.functions.TernaryDense(False), c)
    check(c, res, 5e-5)
    # the +-0.5 / 0 weights of row 0 (kept as +-0.5 / 0 by the torch.sign form) show up in grad_input = g . W_t
    from pytorch_quantize_impls_b200.functions.terner_connect import _functional_ternary
    wt = _functional_ternary(c["w"].cuda(), False).cpu()
    assert wt[0, :6].tolist() == [0.0, 0.0, 0.5, -0.5, 0.0, -1.0]


@pytest.mark.parametrize("tag", ["s1p1", "s2p0", "nobias"])
def test_ternary_conv(Q, fgold, tag):
    c = fcase(fgold, "ternary_conv_" + tag)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        op = Q.functions.TernaryConv2d(False, **CONV_KW[tag])
    check(c, run(op, c), 1e-4, grad_tol=1e-4)       # conv gradients on the tensor-core hi/lo route (engine.grad_*_conv2d)


@pytest.mark.parametrize("name,k", [("quant_dense_k1", 1), ("quant_dense_k2", 2), ("quant_dense_k3", 3), ("quant_dense_k4", 4),
                                    ("quant_dense_k32", 32), ("quant_dense_k3_real", 3)])
def test_quant_dense(Q, fgold, name, k):
    c = fcase(fgold, name)
    check(c, run(Q.functions.QuantDense(k), c), 1e-4 if k == 32 else 5e-5)


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("tag", ["s1p1", "s2p0"])
def test_quant_conv(Q, fgold, k, tag):
    c = fcase(fgold, f"quant_conv_k{k}_{tag}")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        op = Q.functions.QuantConv2d(bit_width=k, **CONV_KW[tag])
    check(c, run(op, c), 1e-4, grad_tol=1e-4)
