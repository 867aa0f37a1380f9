"""Whole-network converters -- the surface of QuantTorch/utils/convertor.py:21-76.

`xxx_net_convert(net, ...)` returns a deep copy of `net` in which every `nn.Linear` / `nn.Conv2d` has been replaced by the
quantized layer of one family (through that layer's own static `convert`, e.g. binary_layers.py:8-12).  Differences from the
reference, all deliberate:

  * the reference's helpers pass keyword names its layers no longer accept (`weight_bit`, `bitwight`, `quant_input` for the
    dense layer: utils/convertor.py:50-69) and fail with TypeError; here the reference's argument names are kept on the
    helper signatures and mapped onto the layers' real parameters (`bit_width`, ...);
  * the layers' `convert` creates freshly initialised layers (as in the reference).  `copy_weights=True` additionally copies
    `weight` / `bias` from the layer being replaced, which is what one wants when quantizing a trained fp32 model;
  * `ternary_net_convert` exists (the reference has ternary layers but no converter for them);
  * the elastic / WQR converters are not provided (training-time regularisers, out of scope: DESIGN.md section 8).
"""
from copy import deepcopy

import torch
from torch import nn

from .layers.binary_layers import BinConv2d, LinearBin
from .layers.dorefa_layers import DorefaConv2d, LinearDorefa
from .layers.log_lin_layers import LinearQuant, QuantConv2d
from .layers.terner_layers import LinearTer, TerConv2d
from .layers.xnor_layers import LinearXNOR, XNORConv2d


def _swap(module, table, copy_weights):
    """Replace, in place and recursively, the children of `module` whose exact class is a key of `table`."""
    for name, child in list(module.named_children()):
        entry = table.get(type(child))
        if entry is None:
            _swap(child, table, copy_weights)
            continue
        target, kwargs = entry
        new = target.convert(child, **kwargs)
        if copy_weights:
            with torch.no_grad():
                new.weight.copy_(child.weight)
                if child.bias is not None and new.bias is not None:
                    new.bias.copy_(child.bias)
        new.to(child.weight.device)          # layers come back in training mode (fp32 master weights), as in the reference:
        setattr(module, name, new)           # call .eval() on the converted net to pack them

    return module


def convert(module, replace_dict, copy_weights=False):
    """Generic form (utils/convertor.py:36-37): replace_dict maps a layer class to `(quantized class, convert kwargs)` or to a
    quantized class alone.  The root module itself is converted when its class is a key."""
    table = {k: (v if isinstance(v, tuple) else (v, {})) for k, v in replace_dict.items()}
    module = deepcopy(module)
    holder = nn.Module()
    holder.root = module
    return _swap(holder, table, copy_weights).root


def binary_net_convert(net, deterministic=True, copy_weights=False):
    kw = {"deterministic": deterministic}
    return convert(net, {nn.Linear: (LinearBin, kw), nn.Conv2d: (BinConv2d, kw)}, copy_weights)


def ternary_net_convert(net, deterministic=True, copy_weights=False):
    kw = {"deterministic": deterministic}
    return convert(net, {nn.Linear: (LinearTer, kw), nn.Conv2d: (TerConv2d, kw)}, copy_weights)


def dorefa_net_convert(net, weight_bit=3, copy_weights=False):
    kw = {"bit_width": weight_bit}
    return convert(net, {nn.Linear: (LinearDorefa, kw), nn.Conv2d: (DorefaConv2d, kw)}, copy_weights)


def xnor_net_convert(net, dim=[0, 1], quant_input=False, copy_weights=False):
    return convert(net, {nn.Linear: (LinearXNOR, {"dim": dim}),
                         nn.Conv2d: (XNORConv2d, {"dim": dim, "quant_input": quant_input})}, copy_weights)


def log_lin_net_convert(net, fsr=7, bitwight=3, dtype="lin", copy_weights=False):
    kw = {"fsr": fsr, "bit_width": bitwight, "dtype": dtype}
    return convert(net, {nn.Linear: (LinearQuant, kw), nn.Conv2d: (QuantConv2d, kw)}, copy_weights)
