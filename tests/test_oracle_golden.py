"""CPU: the oracle restatement reproduces every golden vector (outputs of the live
reference, oracle/gen_golden.py) bit for bit, and the reference's own KATs."""
import numpy as np
import pytest
import torch

import quanttorch_oracle as O
from conftest import case


def eq(a, b):
    return torch.equal(torch.nan_to_num(a, nan=12345.0), torch.nan_to_num(b, nan=12345.0))


def test_elementwise(golden):
    c = case(golden, "safe_sign"); assert eq(O.safe_sign(c["x"]), c["out"])
    c = case(golden, "binary_det"); assert eq(O.binary_det(c["x"]), c["out"])
    c = case(golden, "ternary_det"); assert eq(O.ternary_det(c["x"]), c["out"])


def test_reference_kats(golden):
    # tests/implementations/Terner/function_test.py:10-27
    assert O.ternary_det(torch.tensor([0.75, 0.5, 0.25, 0, -1, -0.2])).tolist() == [1, 1, 0, 0, -1, 0]
    assert O.ternary_det(torch.tensor([1, 0, .51, .1, 0, -1, -.2, .7])).tolist() == [1, 0, 1, 0, 0, -1, 0, 1]
    c = case(golden, "ternary_kat1"); assert c["out"].tolist() == [1, 1, 0, 0, -1, 0]
    # survey-probed vectors
    assert O.safe_sign(torch.tensor([-0.0, 0.0, float("nan"), 1e-45, -1e-45])).tolist() == [1, 1, 1, 1, -1]
    assert O.ternary_det(torch.tensor([.75, .5, .25, 0, -.25, -.5, -.75, -1])).tolist() == [1, 1, 0, 0, 0, 0, -1, -1]
    assert (O.dorefa_quantize(torch.tensor([0.5, 1 / 6, 0.1667, 0.8333, 1.2, -0.3]), 2) * 3).tolist() == [2, 0, 1, 2, 4, -1]
    assert O.log_quant(torch.tensor([0.3, 1.234, 5, -1, 0]), 7, 3).tolist() == [0.5, 1, 4, -1, 0]
    # BinaryNet/layer_test.py:16-21: weight 0 -> +1
    c = case(golden, "lin_bin_kat")
    assert eq(O.linear_bin(c["x"], c["w"]), c["out"]) and c["out"].item() == 2 + 1 + 3
    # Dorefa/function_test.py:278-288 all-zero weight
    assert eq(O.dorefa_weight(torch.zeros(3, 5), 3), torch.zeros(3, 5))


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8, 32])
def test_dorefa_ops(golden, k):
    c = case(golden, f"dorefa_quant_k{k}"); assert eq(O.dorefa_quantize(c["x"], k), c["out"])
    c = case(golden, f"dorefa_weight_k{k}"); assert eq(O.dorefa_weight(c["w"], k), c["out"])
    if k not in (1, 32):
        n = 2 ** k - 1
        codes = O.dorefa_weight_codes(c["w"], k)
        assert codes.min() >= 0 and codes.max() <= n
        recon = (2 * torch.from_numpy(codes).float() - n) / n
        assert torch.allclose(recon, c["out"], atol=2e-7)


@pytest.mark.parametrize("d", [-1, 0, 1])
def test_xnor_act(golden, d):
    c = case(golden, f"xnor_act_dim{d}"); assert eq(O.xnor_act(c["x"], d), c["out"])


@pytest.mark.parametrize("fsr,bw", [(7, 3), (2, 2), (5, 4)])
def test_loglin_ops(golden, fsr, bw):
    c = case(golden, f"log_quant_{fsr}_{bw}"); assert eq(O.log_quant(c["x"], fsr, bw), c["out"])
    c = case(golden, f"lin_quant_{fsr}_{bw}"); assert eq(O.lin_quant(c["x"], fsr, bw), c["out"])


def test_dense_layers(golden):
    c = case(golden, "lin_bin")
    assert eq(O.binary_det(c["x"]), c["xq"])
    assert eq(O.linear_bin(c["xq"], c["w"], c["b"]), c["out"])
    assert eq(O.linear_bin(c["x"], c["w"], c["b"]), c["out_real"])
    c2 = case(golden, "lin_bin_eval")
    assert eq(O.binary_det(c["w"]), c2["w_eval"]) and eq(c2["out"], c["out"])
    c = case(golden, "lin_ter")
    assert eq(O.linear_ter(c["xq"], c["w"], c["b"]), c["out"])
    c = case(golden, "lin_xnor")
    assert eq(O.xnor_act(c["x"], 1), c["xq"])
    assert eq(O.linear_xnor(c["xq"], c["w"], c["b"]), c["out"])
    for k in (1, 2, 3, 4, 8):
        for ka in (k, 8 if k != 8 else 4):
            c = case(golden, f"lin_dorefa_w{k}a{ka}")
            assert eq(O.dorefa_quantize(c["x"], ka), c["xq"])
            assert eq(O.linear_dorefa(c["xq"], c["w"], c["b"], k), c["out"])
    for dt, fsr, bw in (("lin", 7, 3), ("log", 7, 3), ("log", 2, 2)):
        c = case(golden, f"lin_loglin_{dt}_{fsr}_{bw}")
        assert eq(O.linear_loglin(c["x"], c["w"], c["b"], dt, fsr, bw), c["out"])


CONV = {"s1p1": dict(stride=1, padding=1), "s2p0": dict(stride=2, padding=0),
        "s1p2d2": dict(stride=1, padding=2, dilation=2)}


@pytest.mark.parametrize("tag", list(CONV))
def test_conv_layers(golden, tag):
    kw = CONV[tag]
    c = case(golden, f"conv_bin_{tag}")
    assert eq(O.conv_bin(c["xq"], c["w"], c["b"], **kw), c["out"])
    c = case(golden, f"conv_ter_{tag}")
    assert eq(O.conv_ter(c["xq"], c["w"], c["b"], **kw), c["out"])
    for k in (2, 4, 8):
        c = case(golden, f"conv_dorefa_w{k}a{k}_{tag}")
        assert eq(O.conv_dorefa(c["xq"], c["w"], c["b"], k, **kw), c["out"])
    c = case(golden, f"conv_xnor_{tag}")
    assert eq(O.conv_xnor(c["x"], c["w"], c["b"], **kw), c["out"])


def test_integer_identities(golden):
    """The identities the low-bit kernels rely on hold exactly against reference outputs."""
    c = case(golden, "lin_bin")
    K = c["x"].shape[1]
    acc = O.int_acc(O.sign_codes(c["x"]), O.sign_codes(c["w"]))
    pop = O.xnor_popcount_acc(O.sign_bits_packed(c["x"]), O.sign_bits_packed(c["w"]), K)
    assert np.array_equal(acc, pop)
    y = torch.from_numpy(acc).float() + c["b"]
    assert eq(y, c["out"])                         # bit-exact incl. bias
    c = case(golden, "lin_ter")
    acc = O.int_acc(O.sign_codes(c["x"]), O.ternary_codes(c["w"]))
    assert eq(torch.from_numpy(acc).float() + c["b"], c["out"])
    for k in (2, 4, 8):
        c = case(golden, f"lin_dorefa_w{k}a{k}")
        n = 2 ** k - 1
        ca, cw = O.dorefa_act_codes(c["x"], k), O.dorefa_weight_codes(c["w"], k)
        acc = O.int_acc(ca, 2 * cw - n)
        y = torch.from_numpy(acc).double() / (n * n) + c["b"].double()
        rel = (y - c["out"].double()).abs().max() / c["out"].abs().max()
        assert rel < 1e-5, rel
