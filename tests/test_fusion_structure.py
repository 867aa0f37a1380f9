"""CPU: `fuse_inference` is a pure module-graph rewrite -- which runs of modules fold into which fused module can be checked
without a GPU (the fused modules fall back to the plain composition on anything but the inference chain, which is GPU work)."""
import torch
from torch import nn

import pytorch_quantize_impls_b200 as Q
from pytorch_quantize_impls_b200 import nets

F, L = Q.functions, Q.layers


def names(seq):
    return [type(m).__name__ for m in seq]


def test_linear_chain_patterns():
    net = nn.Sequential(F.BinaryConnect(), L.LinearBin(64, 64), nn.BatchNorm1d(64), nn.Hardtanh(), F.BinaryConnect(),
                        L.LinearTer(64, 64), F.nnDorefaQuant(4), L.LinearDorefa(64, 32, bit_width=4), nn.BatchNorm1d(32),
                        nn.ReLU(), L.LinearBin(32, 8))
    fused = Q.fuse_inference(net)
    assert names(fused) == ["fronteur", "FusedLayerQuant", "FusedLayerQuant", "FusedLayerBN", "LinearBin"]
    a, b, c = fused[1], fused[2], fused[3]
    assert isinstance(a.layer, L.LinearBin) and isinstance(a.bn, nn.BatchNorm1d) and isinstance(a.act, nn.Hardtanh)
    assert isinstance(b.layer, L.LinearTer) and b.bn is None and b.act is None
    assert isinstance(c.layer, L.LinearDorefa) and isinstance(c.bn, nn.BatchNorm1d) and isinstance(c.act, nn.ReLU)
    # lane format hint: the sign codes feeding LinearTer may be e2m1, the DoReFa-4 codes feeding LinearDorefa(4) need 8-bit lanes
    assert a._consumer_needs_i8 is False and b._consumer_needs_i8 is True


def test_pool_between_layer_and_quantizer_runs_on_codes():
    net = nn.Sequential(L.DorefaConv2d(3, 32, 3, bit_width=4), nn.MaxPool2d(2), nn.BatchNorm2d(32), nn.Hardtanh(0., 1.),
                        F.nnDorefaQuant(4), L.DorefaConv2d(32, 64, 3, bit_width=4), nn.BatchNorm2d(64), nn.Hardtanh(0., 1.),
                        F.nnDorefaQuant(4), L.DorefaConv2d(64, 64, 3, groups=2, bit_width=4), nn.BatchNorm2d(64))
    fused = Q.fuse_inference(net)
    # conv -> pool -> BN -> clamp -> quantizer: the quantizer runs in the conv epilogue, the pool on the codes
    assert names(fused) == ["FusedLayerPoolQuant", "FusedLayerQuant", "DorefaConv2d", "BatchNorm2d"]
    assert isinstance(fused[0].pool, nn.MaxPool2d) and isinstance(fused[0].bn, nn.BatchNorm2d)


def test_resnet_blocks_and_unknown_modules_are_left_alone():
    net = Q.fuse_inference(nets.resnet18_ternary())
    assert names(net.stem) == ["FusedConvPool"] and isinstance(net.stem[0].inner.act, nn.Hardtanh)
    assert set(names(net.layers)) == {"FusedBasicBlock"}
    assert net.layers[0]._next == ("dorefa", 8) and net.layers[-1]._next is None
    # blocks in front of a down-sampling block (conv shortcut) hand over codes only
    assert [b._next_reads_fp32 for b in net.layers] == [True, False, True, False, True, False, True, True]
    blk = net.layers[2].block                 # first down-sampling block
    assert names(blk.branch1) == ["FusedLayerQuant"] and names(blk.branch2) == ["FusedLayerBN"]
    assert names(blk.shortcut) == ["FusedLayerBN"] and isinstance(net.linear, nn.Linear)
    assert type(net.__dict__["_fused_head"]).__name__ == "FusedAvgLinear" and "_fused_head" not in dict(net.named_children())
    plain = nn.Sequential(nn.Linear(4, 4), nn.BatchNorm1d(4), nn.ReLU())
    assert names(Q.fuse_inference(plain)) == ["Linear", "BatchNorm1d", "ReLU"]          # not a quantized layer: untouched


def test_fused_modules_are_the_plain_composition_in_training_mode():
    """With autograd on (or BatchNorm in training mode) the fused wrappers must run their children one by one; on the CPU that
    reaches the first quantized child, which refuses CPU tensors loudly (no CPU fallback)."""
    import pytest
    net = Q.fuse_inference(nn.Sequential(L.LinearBin(8, 8), nn.BatchNorm1d(8), nn.Hardtanh(), F.BinaryConnect()))
    assert names(net) == ["FusedLayerQuant"]
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        net(torch.randn(4, 8))


def test_network_configs_fuse_into_code_chains():
    """BASELINE configs 3 and 5: every conv block folds its BatchNorm + clamp + quantizer (and pool) into the conv, the classifier
    reads the channels-last flatten of the codes, no stand-alone BatchNorm / pool / quantizer is left."""
    for net in (nets.alexnet_dorefa(bit_width=4), nets.vgg_dorefa(bit_width=8)):
        fused = names(Q.fuse_inference(net))
        assert not ({"BatchNorm2d", "BatchNorm1d", "MaxPool2d", "Hardtanh", "fronteur", "Flatten"} & set(fused)), fused
        assert fused.count("FlattenCodes") == 1 and fused[-1] == "LinearDorefa"
    head = Q.fuse_inference(nn.Sequential(F.BinaryConnect(), L.LinearBin(4096, 4096)))
    assert names(head) == ["FusedActLayer"]


def test_resnet_head_falls_back_to_the_two_modules_off_the_gpu():
    """FusedAvgLinear is one batch-invariant kernel on CUDA inference; anywhere else it must be avg-pool -> flatten -> Linear."""
    torch.manual_seed(0)
    net = Q.fuse_inference(nets.resnet18_ternary())
    head = net.__dict__["_fused_head"]
    x = torch.randn(3, 512, 4, 5)
    with torch.no_grad():
        assert torch.equal(head(x), net.linear(net.avg(x).flatten(1)))
    assert "_fused_head" not in net.state_dict() and all(not k.startswith("_fused_head") for k in net.state_dict())


def test_overflow_flag_policy():
    """The sticky lane-overflow flag exists only where somebody can look at it (ops.set_strict)."""
    from pytorch_quantize_impls_b200 import _lib as lib
    from pytorch_quantize_impls_b200 import _ops as ops
    dev = torch.device("cpu")
    try:
        ops.set_strict(False)
        assert ops._overflow_flag(dev, lib.Q_DOREFA) is not None and ops._overflow_flag(dev, lib.Q_SIGN) is None
        assert ops._overflow_flag(dev, lib.Q_DOREFA, guaranteed=True) is None
        ops.set_strict("off")
        assert ops._overflow_flag(dev, lib.Q_DOREFA) is None
    finally:
        ops.set_strict(False)
    assert ops.clamp_guarantees_lane(lib.CODES_U8, 8, 0.0, 1.0) and not ops.clamp_guarantees_lane(lib.CODES_I8, 8, 0.0, 1.0)
