"""CPU, build container only: the oracle equals the LIVE reference bit for bit on seeded
random tensors at non-fixture shapes.  Skipped where /root/reference is absent (GPU box)."""
import pytest
import torch

import quanttorch_oracle as O
from ref_loader import load_reference

REF = load_reference()
pytestmark = pytest.mark.skipif(REF is None, reason="reference tree not present")


@pytest.mark.parametrize("seed", [0, 1])
def test_ops_and_layers_match_live_reference(seed):
    Fn, L = REF
    torch.manual_seed(seed)
    x = torch.randn(17, 50) * 1.3
    xu = torch.rand(17, 50)
    w = torch.randn(11, 50) * 0.4
    b = torch.rand(11) - 0.5
    assert torch.equal(O.binary_det(x), Fn.BinaryConnectDeterministic.apply(x))
    assert torch.equal(O.ternary_det(x), Fn.TernaryConnectDeterministic.apply(x))
    for k in (1, 2, 3, 4, 6, 8):
        assert torch.equal(O.dorefa_quantize(xu, k), Fn.DorefaQuant(xu, k))
        assert torch.equal(O.dorefa_weight(w, k), Fn.nnQuantWeight(k)(w))
        lay = L.LinearDorefa(50, 11, bit_width=k); lay.weight.data.copy_(w); lay.bias.data.copy_(b)
        xq = Fn.DorefaQuant(xu, k)
        assert torch.equal(O.linear_dorefa(xq, w, b, k), lay(xq))
    for d in (-1, 0, 1):
        assert torch.equal(O.xnor_act(x, d), Fn.QuantXnor(x, d))
    lay = L.LinearXNOR(50, 11); lay.weight.data.copy_(w); lay.bias.data.copy_(b)
    assert torch.equal(O.linear_xnor(x, w, b), lay(x))
    lay = L.LinearBin(50, 11); lay.weight.data.copy_(w); lay.bias.data.copy_(b)
    assert torch.equal(O.linear_bin(x, w, b), lay(x))
    lay = L.LinearTer(50, 11); lay.weight.data.copy_(w); lay.bias.data.copy_(b)
    assert torch.equal(O.linear_ter(x, w, b), lay(x))
    xi = torch.randn(2, 4, 8, 8); wc = torch.randn(6, 4, 3, 3) * 0.5; bc = torch.rand(6)
    for cls, fn in ((L.BinConv2d, O.conv_bin), (L.TerConv2d, O.conv_ter), (L.XNORConv2d, O.conv_xnor)):
        lay = cls(4, 6, 3, stride=2, padding=1); lay.weight.data.copy_(wc); lay.bias.data.copy_(bc)
        assert torch.equal(fn(xi, wc, bc, stride=2, padding=1), lay(xi))
    for dt in ("lin", "log"):
        assert torch.equal(O.loglin_weight(w * 50, dt, 5, 3), Fn.Quant(w * 50, dt, 5, 3))
