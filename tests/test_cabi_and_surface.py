"""CPU: the C-ABI library builds, loads and exports every symbol include/qtb200.h declares (no compute calls),
and the Python host surface mirrors the reference's names, signatures and error behaviour."""
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def Q():
    import __graft_entry__ as g
    g.build()
    import pytorch_quantize_impls_b200 as Q
    return Q


def test_library_exports_every_declared_symbol(Q):
    from pytorch_quantize_impls_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "qtb200.h")).read()
    declared = set(re.findall(r"\b(qt_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    h = _lib.lib()
    for name in declared:
        assert hasattr(h, name), name
    assert h.qt_version() >= 101
    assert h.qt_launch_count(0) == 0


def test_ctypes_structs_match_header_field_order(Q):
    from pytorch_quantize_impls_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "qtb200.h")).read()
    for cls in (_lib.QtActQuant, _lib.QtWeightPack, _lib.QtWeightExpand, _lib.QtIm2col, _lib.QtEpilogue):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cls.__name__, cls.__name__), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        assert names == [f[0] for f in cls._fields_], cls.__name__


def test_surface_names(Q):
    fn = ["safeSign", "BinaryConnectDeterministic", "BinaryConnectStochastic", "BinaryConnect", "BinaryDense",
          "BinaryConv2d", "nnDorefaQuant", "DorefaQuant", "nnQuantWeight", "QuantDense",
          "QuantConv2d", "LogQuant", "LinQuant", "nnQuant", "Quant", "TernaryConnectDeterministic",
          "TernaryConnectStochastic", "TernaryConnect", "TernaryDense", "TernaryConv2d", "nnQuantXnor", "QuantXnor",
          "XNORDense", "XNORConv2d"]
    for n in fn:
        assert hasattr(Q.functions, n), n
    for n in ["LinearBin", "BinConv2d", "LinearDorefa", "DorefaConv2d",
              "LinearQuant", "QuantConv2d", "LinearTer", "TerConv2d", "LinearXNOR", "XNORConv2d"]:
        assert hasattr(Q.layers, n), n
    assert Q.BinaryNet.LinearBin is Q.layers.LinearBin and Q.BinaryNet.BinaryConnect is Q.functions.BinaryConnect
    assert Q.XnorNet.LinearXNOR is Q.layers.LinearXNOR and Q.DorefaNet.DorefaConv2d is Q.layers.DorefaConv2d
    assert Q.TernerNet.TerConv2d is Q.layers.TerConv2d and Q.LogLinNet.LinearQuant is Q.layers.LinearQuant


def test_signatures_match_reference(Q):
    L = Q.layers
    def params(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values() if p.name != "self"]
    E = inspect.Parameter.empty
    assert params(L.LinearBin.__init__) == [("in_features", E), ("out_features", E), ("bias", True), ("deterministic", True)]
    assert params(L.BinConv2d.__init__)[-6:] == [("stride", 1), ("padding", 0), ("dilation", 1), ("groups", 1),
                                                  ("bias", True), ("deterministic", True)]
    assert params(L.LinearDorefa.__init__)[-1] == ("bit_width", 3)
    assert params(L.LinearQuant.__init__)[-3:] == [("dtype", "lin"), ("fsr", 7), ("bit_width", 3)]
    assert params(L.QuantConv2d.__init__)[-3:] == [("fsr", 7), ("bit_width", 3), ("dtype", "lin")]
    assert params(L.LinearXNOR.__init__)[-1] == ("dim", [0, 1])
    assert params(Q.functions.DorefaQuant) == [("x", E), ("bit_width", 3)]
    assert params(Q.functions.QuantXnor) == [("input", E), ("dim", 1)]
    assert params(Q.functions.Quant)[1:] == [("dtype", "lin"), ("fsr", 7), ("bit_width", 3), ("with_sign", True), ("lin_back", True)]


def test_module_contract_on_cpu(Q):
    L = Q.layers
    lay = L.LinearBin(16, 4)
    assert isinstance(lay, torch.nn.Linear) and isinstance(lay, L.QLayer)
    assert list(lay.state_dict().keys()) == ["weight", "bias"]
    assert float(lay.bias.abs().sum()) == 0.0                       # reset_parameters: bias zero, binary_layers.py:20-23
    lay.weight.data.fill_(3.0); lay.clamp(); assert float(lay.weight.max()) == 1.0
    for cls, other in ((L.LinearBin, torch.nn.Conv2d(1, 1, 1)), (L.BinConv2d, torch.nn.Linear(1, 1)),
                       (L.LinearTer, torch.nn.Conv2d(1, 1, 1)), (L.DorefaConv2d, torch.nn.Linear(1, 1)),
                       (L.LinearXNOR, torch.nn.Conv2d(1, 1, 1)), (L.LinearQuant, torch.nn.Conv2d(1, 1, 1))):
        with pytest.raises(TypeError):
            cls.convert(other)
    c = L.BinConv2d.convert(torch.nn.Conv2d(3, 8, 5, stride=2, padding=1, bias=False))
    assert c.kernel_size == (5, 5) and c.stride == (2, 2) and c.bias is None
    d = L.LinearDorefa.convert(torch.nn.Linear(5, 6), bit_width=4)
    assert d.bit_width == 4 and "bit_width = 4" in repr(d)
    q = L.LinearQuant(8, 8, fsr=5, bit_width=2)
    assert float(q.weight.abs().min()) >= 2 ** 3 and float(q.weight.abs().max()) <= 2 ** 5
    with pytest.raises(RuntimeError):
        Q.functions.nnQuantXnor(2)
    with pytest.raises(RuntimeError):
        Q.functions.Quant(torch.zeros(1), dtype="exp")
    with pytest.raises(NotImplementedError):
        lay.get_quant_weight()


def test_no_cpu_fallback(Q):
    """CPU tensors are refused loudly; nothing routes through torch CPU math or the oracle."""
    lay = Q.layers.LinearBin(8, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        lay(torch.zeros(2, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        Q.functions.DorefaQuant(torch.zeros(2, 8), 4)
    import pytorch_quantize_impls_b200, sys
    src_dir = os.path.dirname(pytorch_quantize_impls_b200.__file__)
    for root, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(root, f)).read()
                assert "quanttorch_oracle" not in txt and "import oracle" not in txt, f


def test_ctypes_mirrors_have_the_c_struct_sizes():
    """Every ctypes.Structure in _lib.py must have the size the compiler gives the struct of include/qtb200.h."""
    import ctypes
    from pytorch_quantize_impls_b200 import _lib as L
    lib = L.lib()
    for name in ("QtActQuant", "QtWeightPack", "QtWeightExpand", "QtIm2col", "QtRequant", "QtEpilogue", "QtConvGeom"):
        assert lib.qt_sizeof(name.encode()) == ctypes.sizeof(getattr(L, name)), name
    assert lib.qt_sizeof(b"nope") == -1


@pytest.mark.parametrize("mk", [
    lambda Q: Q.layers.LinearBin(8, 4), lambda Q: Q.layers.LinearBin(8, 4, deterministic=False),
    lambda Q: Q.layers.LinearTer(8, 4), lambda Q: Q.layers.TerConv2d(2, 4, 3, deterministic=False),
    lambda Q: Q.layers.LinearDorefa(8, 4, bit_width=3), lambda Q: Q.layers.DorefaConv2d(2, 4, 3, bit_width=1),
    lambda Q: Q.layers.LinearQuant(8, 4, dtype="log"), lambda Q: Q.layers.LinearQuant(8, 4, dtype="lin"),
    lambda Q: Q.layers.LinearXNOR(8, 4)])
def test_eval_swap_before_moving_to_the_gpu(mk):
    """`model.eval(); model.cuda()` is a common drop-in order and the reference supports eval() on CPU weights
    (binary_layers.py:30-40): the swap happens with host arithmetic, packing waits for the first CUDA forward, and a failed
    swap never leaves the layer half-switched."""
    import pytorch_quantize_impls_b200 as Q
    import quanttorch_oracle as O
    torch.manual_seed(5)
    lay = mk(Q)
    w0 = lay.weight.data.clone()
    lay.eval()
    assert lay.training is False and torch.equal(lay.weight.org, w0)
    wq = lay.weight.data
    name = type(lay).__name__
    if getattr(lay, "deterministic", True):
        ref = {"LinearBin": lambda: O.binary_det(w0), "LinearTer": lambda: O.ternary_det(w0),
               "LinearDorefa": lambda: O.dorefa_weight(w0, 3), "DorefaConv2d": lambda: O.dorefa_weight(w0, 1),
               "LinearQuant": lambda: O.loglin_weight(w0, lay._dtype, lay.fsr, lay.bit_width),
               "LinearXNOR": lambda: O.xnor_weight(w0)}[name]()
        assert torch.equal(wq, ref)
    else:
        assert set(wq.unique().tolist()) <= {-1.0, 0.0, 1.0}
    lay.eval()                                   # idempotent
    assert torch.equal(lay.weight.data, wq)
    lay.train()
    assert lay.training is True and torch.equal(lay.weight.data, w0)
