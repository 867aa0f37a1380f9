"""Packed-weight checkpoints: the k-bit HBM format of every quantized layer on disk.

The reference checkpoints fp32 master weights (`torch.save(model.state_dict())`, train/early_stopping.py:47-54) and
re-quantizes them on load.  `save_packed` stores what the kernels actually read -- sign bits, ternary / XnorNet bit planes
(+ alpha[k]), DoReFa k-bit codes (+ their per-tensor scale) -- next to the ordinary state of every other module, and
`load_packed` installs those packs straight into the layers: no fp32 master copy exists afterwards (a 4096 x 4096 binary
layer is 2 MB instead of 67 MB), the layers are inference-only until fp32 weights are loaded again.

File format ("qtb200-packed-v1", a torch.save'd dict):
    {"format": ..., "layers": {module_name: {kind, bit_width, n, k, ld_packed, weight_shape, packed (uint8 [planes, n, ld]),
                                              alpha, alpha_norm, alpha_max, stats, col_scale, planes, ld_planes}},
     "state": {every state_dict entry except the quantized layers' `weight`}}
Bit layout of `packed`: include/qtb200.h (QtWeightPack).
"""
import torch

from . import _ops as ops
from .layers.common import QuantLayerMixin

FORMAT = "qtb200-packed-v1"
_TENSORS = ("packed", "alpha", "alpha_norm", "alpha_max", "stats", "col_scale", "planes")
_SCALARS = ("kind", "bit_width", "n", "k", "ld_packed", "ld_planes", "wscale", "emin")


def _quant_layers(model):
    return [(name, m) for name, m in model.named_modules() if isinstance(m, QuantLayerMixin)]


def packed_state(model):
    """Packed state of `model` (tensors on the CPU).  Layers are packed from their fp32 master weights (eval-mode layers
    reuse the pack made at train(False)); packed-only layers re-export the pack they hold."""
    layers = {}
    skip = set()
    for name, m in _quant_layers(model):
        pack = m._current_pack()
        ent = {k: getattr(pack, k) for k in _SCALARS}
        for k in _TENSORS:
            t = getattr(pack, k)
            ent[k] = None if t is None else t.detach().cpu()
        ent["weight_shape"] = tuple(m._wshape())
        layers[name] = ent
        skip.add((name + "." if name else "") + "weight")
    state = {k: v.detach().cpu() for k, v in model.state_dict().items() if k not in skip}
    return {"format": FORMAT, "layers": layers, "state": state}


def save_packed(model, path):
    torch.save(packed_state(model), path)


def load_packed(model, state, drop_master=True):
    """Install a packed state (dict or path) into `model` (same architecture, already on its device).  Every quantized
    layer becomes packed-only and inference-only; with drop_master=True its fp32 `weight` storage is released."""
    if not isinstance(state, dict):
        state = torch.load(state, map_location="cpu")
    if state.get("format") != FORMAT:
        raise ValueError("not a %s checkpoint (format = %r)" % (FORMAT, state.get("format")))
    layers = dict(_quant_layers(model))
    missing = set(layers) - set(state["layers"])
    extra = set(state["layers"]) - set(layers)
    if missing or extra:
        raise KeyError("packed checkpoint does not match the model: missing %s, unexpected %s" % (sorted(missing), sorted(extra)))
    for name, m in layers.items():
        ent = state["layers"][name]
        if tuple(ent["weight_shape"]) != tuple(m._wshape()):
            raise ValueError("layer %s: checkpoint weight shape %s, model %s" % (name, tuple(ent["weight_shape"]), tuple(m._wshape())))
        dev = m.weight.device if m.weight.numel() else (m.bias.device if m.bias is not None else torch.device("cuda"))
        ops.require_cuda(torch.empty(0, device=dev), "model (load_packed installs device-resident packs)")
        pack = ops.WeightPack()
        for k in _SCALARS:
            setattr(pack, k, ent[k])
        for k in _TENSORS:
            t = ent[k]
            setattr(pack, k, None if t is None else t.to(dev).contiguous())
        pack.wq = None
        m._install_packed(pack, tuple(ent["weight_shape"]), drop_master)
    own = model.state_dict()
    rest = {k: v for k, v in state["state"].items() if k in own}
    model.load_state_dict(rest, strict=False)
    return model
