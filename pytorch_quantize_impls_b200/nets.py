"""Network topologies of the BASELINE configs, built on the drop-in layers.

The reference's model zoo (`models/**`) is bit-rotted against its own layer API, so
the named topologies are re-stated here; the quantized layers/quantizers come from `lib` (this package by default;
the tests pass an oracle-backed namespace with the same class names to get a CPU twin of the same graph).

Convention used by every net here: the activation quantizer sits immediately in front of the quantized layer it feeds
(`conv -> pool -> BN -> act -> quantizer -> conv`, as in models/Alexnet/Alexnet_Bin.py:12-54), so the low-bit operand
reaches the next layer without an fp32 detour.  DoReFa k-bit activations are clamped to [0, 1] by Hardtanh(0, 1)
before quantisation (the DoReFa contract, dorefa_connect.py:14-18).
"""
import torch
from torch import nn


def _default_lib():
    import pytorch_quantize_impls_b200 as Q

    class _Lib:
        pass
    lib = _Lib()
    for name in ("LinearBin", "BinConv2d", "LinearXNOR", "XNORConv2d", "LinearTer", "TerConv2d", "LinearDorefa",
                 "DorefaConv2d", "LinearQuant", "QuantConv2d"):
        setattr(lib, name, getattr(Q.layers, name))
    for name in ("BinaryConnect", "TernaryConnect", "nnDorefaQuant", "nnQuantXnor"):
        setattr(lib, name, getattr(Q.functions, name))
    return lib


def xnor_mlp(dims=(4096, 4096, 4096, 1000), lib=None):
    """BASELINE configs[1]: XnorNet MLP, nnQuantXnor(1) -> LinearXNOR per layer."""
    lib = lib or _default_lib()
    mods = []
    for i in range(len(dims) - 1):
        mods += [lib.nnQuantXnor(1), lib.LinearXNOR(dims[i], dims[i + 1])]
    return nn.Sequential(*mods)


def binary_mlp(dims=(4096, 4096), lib=None):
    """North-star layer stack: BinaryConnect() -> LinearBin."""
    lib = lib or _default_lib()
    mods = []
    for i in range(len(dims) - 1):
        mods += [lib.BinaryConnect(), lib.LinearBin(dims[i], dims[i + 1])]
    return nn.Sequential(*mods)


class Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.size(0), -1)


def alexnet_dorefa(bit_width=4, act_bits=None, num_classes=10, coef=3, lib=None):
    """BASELINE configs[2]: AlexNet of models/Alexnet/Alexnet_Bin.py:12-54 (ImageNet shapes, widths x coef) with the
    binary layers swapped for DorefaConv2d / LinearDorefa(bit_width) and nnDorefaQuant(act_bits) activations.
    The first conv sees the fp32 image (real-activation route)."""
    lib = lib or _default_lib()
    k = bit_width
    a = act_bits or bit_width

    def act(ch, dim2=True):
        return [nn.BatchNorm2d(ch) if dim2 else nn.BatchNorm1d(ch), nn.Hardtanh(0.0, 1.0), lib.nnDorefaQuant(a)]

    features = [
        lib.DorefaConv2d(3, 64 * coef, kernel_size=11, stride=4, padding=2, bit_width=k),
        nn.MaxPool2d(kernel_size=3, stride=2), *act(64 * coef),
        lib.DorefaConv2d(64 * coef, 192 * coef, kernel_size=5, padding=2, bit_width=k),
        nn.MaxPool2d(kernel_size=3, stride=2), *act(192 * coef),
        lib.DorefaConv2d(192 * coef, 384 * coef, kernel_size=3, padding=1, bit_width=k), *act(384 * coef),
        lib.DorefaConv2d(384 * coef, 256 * coef, kernel_size=3, padding=1, bit_width=k), *act(256 * coef),
        lib.DorefaConv2d(256 * coef, 256, kernel_size=3, padding=1, bit_width=k),
        nn.MaxPool2d(kernel_size=3, stride=2), *act(256),
    ]
    # the quantizer sits in front of Flatten (elementwise: the same values either way), so that the conv chain hands its
    # channels-last codes straight to the classifier (fusion.FlattenCodes)
    classifier = [
        Flatten(),
        lib.LinearDorefa(256 * 6 * 6, 4096, bit_width=k), *act(4096, False),
        lib.LinearDorefa(4096, 4096, bit_width=k), *act(4096, False),
        lib.LinearDorefa(4096, num_classes, bit_width=k),
    ]
    return nn.Sequential(*features, *classifier)


class TerBasicBlock(nn.Module):
    """BasicBlock of models/Resnet/Resnet_bin.py:7-32 with ternary weights and k-bit activations; the reference's
    `conv2(x)` typo (:29, feeds x instead of out) is fixed."""
    expansion = 1

    def __init__(self, lib, in_planes, planes, stride, act_bits):
        super().__init__()
        self.q_in = lib.nnDorefaQuant(act_bits)
        # conv -> BN -> clamp -> quantizer kept as one Sequential so that fusion.fuse_inference can move BN, clamp and the
        # quantizer into the conv epilogue (the codes go straight to conv2's TMA im2col)
        self.branch1 = nn.Sequential(lib.TerConv2d(in_planes, planes, kernel_size=3, stride=stride, padding=1, bias=False),
                                     nn.BatchNorm2d(planes), nn.Hardtanh(0.0, 1.0), lib.nnDorefaQuant(act_bits))
        # conv -> BN pairs are Sequentials so that fuse_inference folds the BatchNorm into the conv epilogue (FusedLayerBN)
        self.branch2 = nn.Sequential(lib.TerConv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False),
                                     nn.BatchNorm2d(planes))
        self.clip = nn.Hardtanh(0.0, 1.0)
        self.shortcut = None
        if stride != 1 or in_planes != planes:
            self.shortcut = nn.Sequential(
                lib.TerConv2d(in_planes, planes, kernel_size=1, stride=stride, bias=False), nn.BatchNorm2d(planes))

    def forward(self, x):                      # x in [0, 1]
        xq = self.q_in(x)
        out = self.branch1(xq)
        out = self.branch2(out)
        out = out + (x if self.shortcut is None else self.shortcut(xq))
        return self.clip(out)


class ResNetTer(nn.Module):
    _uses_fused_head = True        # forward() calls the head fuse_inference leaves in __dict__["_fused_head"]

    def __init__(self, lib, num_blocks, num_classes, act_bits):
        super().__init__()
        self.in_planes = 64
        # ImageNet stem (7x7 s2 + max-pool): the reference's CIFAR stem at 224x224 would cost 27 GMAC/img
        self.stem = nn.Sequential(lib.TerConv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False), nn.BatchNorm2d(64),
                                  nn.Hardtanh(0.0, 1.0), nn.MaxPool2d(kernel_size=3, stride=2, padding=1))
        cfg = [(64, 1), (128, 2), (256, 2), (512, 2)]
        layers = []
        for (planes, stride), n in zip(cfg, num_blocks):
            for s in [stride] + [1] * (n - 1):
                layers.append(TerBasicBlock(lib, self.in_planes, planes, s, act_bits))
                self.in_planes = planes
        self.layers = nn.Sequential(*layers)
        self.avg = nn.AdaptiveAvgPool2d(1)
        self.linear = nn.Linear(512, num_classes)

    def forward(self, x):
        out = self.layers(self.stem(x))
        head = self.__dict__.get("_fused_head")        # set by fuse_inference (not a registered child: same state_dict keys)
        return head(out) if head is not None else self.linear(self.avg(out).flatten(1))


def resnet18_ternary(act_bits=8, num_classes=10, lib=None):
    """BASELINE configs[3]: ResNet-18, ternary (2-bit) weights, 8-bit activations, 224x224."""
    return ResNetTer(lib or _default_lib(), [2, 2, 2, 2], num_classes, act_bits)


def vgg_dorefa(bit_width=8, num_classes=10, lib=None):
    """BASELINE configs[4]: the reference VGG topology (models/VGG/VGG_LinQuant.py:11-61, 32x32 input, 6 convs + 3 fc)
    with DoReFa k-bit weights and activations."""
    lib = lib or _default_lib()
    k = bit_width

    def block(cin, cout, pool, quant=True):
        # conv -> [pool] -> BN -> clamp -> quantizer: the ordering of models/Alexnet/Alexnet_Bin.py (pool right after
        # the conv), which keeps BN/clamp/quantizer adjacent (one fused pass under fusion.fuse_inference)
        m = [lib.DorefaConv2d(cin, cout, kernel_size=3, padding=1, bit_width=k)]
        if pool:
            m.append(nn.MaxPool2d(kernel_size=2))
        m += [nn.BatchNorm2d(cout), nn.Hardtanh(0.0, 1.0)]
        if quant:
            m.append(lib.nnDorefaQuant(k))
        return m

    features = [*block(3, 64, False), *block(64, 64, True), *block(64, 128, False), *block(128, 128, True),
                *block(128, 256, False), *block(256, 256, True)]
    classifier = [Flatten(),                  # quantizer in front of Flatten: see alexnet_dorefa
                  lib.LinearDorefa(4096, 1024, bit_width=k), nn.BatchNorm1d(1024), nn.Hardtanh(0.0, 1.0), lib.nnDorefaQuant(k),
                  lib.LinearDorefa(1024, 1024, bit_width=k), nn.BatchNorm1d(1024), nn.Hardtanh(0.0, 1.0), lib.nnDorefaQuant(k),
                  lib.LinearDorefa(1024, num_classes, bit_width=k)]
    return nn.Sequential(*features, *classifier)
