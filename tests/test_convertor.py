"""CPU: whole-network converters (surface of QuantTorch/utils/convertor.py:21-76) -- pure module-graph rewriting."""
import pytest
import torch
from torch import nn

import pytorch_quantize_impls_b200 as Q
from pytorch_quantize_impls_b200 import convertor as C

L = Q.layers


def fp32_net():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU(),
                         nn.Sequential(nn.Conv2d(8, 8, 3, groups=2, dilation=2), nn.ReLU()), nn.Flatten(),
                         nn.Linear(32, 16), nn.ReLU(), nn.Linear(16, 4, bias=False))


@pytest.mark.parametrize("fn,kw,lin,conv,check", [
    (C.binary_net_convert, dict(deterministic=False), L.LinearBin, L.BinConv2d, lambda m: m.deterministic is False),
    (C.ternary_net_convert, {}, L.LinearTer, L.TerConv2d, lambda m: m.deterministic is True),
    (C.dorefa_net_convert, dict(weight_bit=5), L.LinearDorefa, L.DorefaConv2d, lambda m: m.bit_width == 5),
    (C.xnor_net_convert, dict(dim=[0, 1]), L.LinearXNOR, L.XNORConv2d, lambda m: m.dim == [0, 1]),
    (C.log_lin_net_convert, dict(fsr=2, bitwight=4, dtype="log"), L.LinearQuant, L.QuantConv2d,
     lambda m: (m.fsr, m.bit_width, m._dtype) == (2, 4, "log")),
])
def test_family_converters(fn, kw, lin, conv, check):
    src = fp32_net()
    out = fn(src, **kw)
    assert type(src[0]) is nn.Conv2d and type(src[5]) is nn.Linear                 # the source is untouched (deep copy)
    assert type(out[0]) is conv and type(out[3][0]) is conv and type(out[5]) is lin and type(out[7]) is lin
    assert type(out[1]) is nn.BatchNorm2d and type(out[2]) is nn.ReLU
    c0, c1 = out[0], out[3][0]
    assert (c0.in_channels, c0.out_channels, c0.kernel_size, c0.stride, c0.padding, c0.bias) == (3, 8, (3, 3), (2, 2), (1, 1), None)
    assert (c1.groups, c1.dilation) == (2, (2, 2)) and c1.bias is not None
    assert (out[5].in_features, out[5].out_features) == (32, 16) and out[7].bias is None
    for m in (c0, c1, out[5], out[7]):
        assert check(m)
    assert set(out.state_dict()) == set(src.state_dict())                           # drop-in state_dict keys


def test_copy_weights_and_generic_convert():
    src = fp32_net()
    out = C.binary_net_convert(src, copy_weights=True)
    assert torch.equal(out[0].weight, src[0].weight) and torch.equal(out[5].bias, src[5].bias)
    assert out[0].weight.data_ptr() != src[0].weight.data_ptr()
    fresh = C.binary_net_convert(src)                                               # reference behaviour: no weight copy
    assert not torch.equal(fresh[5].weight, src[5].weight)
    only_fc = C.convert(src, {nn.Linear: (L.LinearDorefa, {"bit_width": 2})})
    assert type(only_fc[0]) is nn.Conv2d and type(only_fc[5]) is L.LinearDorefa and only_fc[5].bit_width == 2
    root = C.convert(nn.Linear(4, 4), {nn.Linear: L.LinearTer})                     # the root itself, class given without kwargs
    assert type(root) is L.LinearTer
    # subclasses are not touched (exact class match, as in the reference): an already quantized layer stays what it is
    again = C.dorefa_net_convert(out, weight_bit=4)
    assert type(again[5]) is L.LinearBin
