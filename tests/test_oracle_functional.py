"""The oracle's restatement of the functional dense / conv ops against golden vectors of the
live reference (oracle/gen_golden_functional.py): forward outputs and the hand-written backward, on the CPU."""
import os

import numpy as np
import pytest
import torch

import quanttorch_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fgold():
    z = np.load(os.path.join(ROOT, "tests", "golden", "quanttorch_ref_functional_v1.npz"))
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def fcase(g, name):
    pre = name + "/"
    return {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}


CONV_KW = {"s1p1": dict(stride=1, padding=1), "s2p0": dict(stride=2, padding=0), "nobias": dict(stride=1, padding=1)}


def _check(c, out, wq, grads, conv=False):
    assert torch.equal(out, c["out"])
    gx, gw, gb = grads
    assert torch.equal(gx, c["gx"])
    if conv:      # oneDNN's weight-gradient summation order depends on the thread count (goldens: 1 thread)
        assert torch.allclose(gw, c["gw"], rtol=1e-5, atol=1e-5)
    else:
        assert torch.equal(gw, c["gw"])
    if "gb" in c:
        assert torch.equal(gb, c["gb"])


@pytest.mark.parametrize("name", ["binary_dense", "binary_dense_nobias", "binary_dense_quant_in"])
def test_binary_dense(fgold, name):
    c = fcase(fgold, name)
    out, wq = O.binary_dense(c["x"], c["w"], c.get("b"))
    _check(c, out, wq, O.dense_grads(c["go"], c["x"], wq, "b" in c))


@pytest.mark.parametrize("name", ["ternary_dense", "ternary_dense_nobias"])
def test_ternary_dense_torch_sign_semantics(fgold, name):
    c = fcase(fgold, name)
    out, wq = O.ternary_dense(c["x"], c["w"], c.get("b"))
    # w[0, :6] = [0, -0, .5, -.5, .25, -.75]: the functional form keeps +-0.5 (terner_connect.py:90), the layer form does not
    assert wq[0, :6].tolist() == [0.0, 0.0, 0.5, -0.5, 0.0, -1.0]
    assert O.ternary_det(c["w"])[0, :6].tolist() == [0.0, 0.0, 1.0, 0.0, 0.0, -1.0]
    _check(c, out, wq, O.dense_grads(c["go"], c["x"], wq, "b" in c))


@pytest.mark.parametrize("tag", ["s1p1", "s2p0", "nobias"])
def test_ternary_conv(fgold, tag):
    c = fcase(fgold, "ternary_conv_" + tag)
    out, wq = O.ternary_conv2d(c["x"], c["w"], c.get("b"), **CONV_KW[tag])
    _check(c, out, wq, O.conv_grads(c["go"], c["x"], c["w"].shape, wq, "b" in c, **CONV_KW[tag]), conv=True)


@pytest.mark.parametrize("name,k", [("quant_dense_k1", 1), ("quant_dense_k2", 2), ("quant_dense_k3", 3), ("quant_dense_k4", 4),
                                    ("quant_dense_k32", 32), ("quant_dense_k3_real", 3)])
def test_quant_dense(fgold, name, k):
    c = fcase(fgold, name)
    out, wq = O.quant_dense(c["x"], c["w"], c.get("b"), bit_width=k)
    gx, gw, gb = O.dense_grads(c["go"], c["x"], wq, "b" in c)
    _check(c, out, wq, (gx, O.dorefa_functional_grad_weight(gw, c["w"], k, conv=False), gb))


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("tag", ["s1p1", "s2p0"])
def test_quant_conv(fgold, k, tag):
    c = fcase(fgold, f"quant_conv_k{k}_{tag}")
    out, wq = O.quant_conv2d(c["x"], c["w"], c["b"], bit_width=k, **CONV_KW[tag])
    gx, gw, gb = O.conv_grads(c["go"], c["x"], c["w"].shape, wq, True, **CONV_KW[tag])
    _check(c, out, wq, (gx, O.dorefa_functional_grad_weight(gw, c["w"], k, conv=True), gb), conv=True)
