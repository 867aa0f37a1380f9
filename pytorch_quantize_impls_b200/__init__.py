"""pytorch_quantize_impls_b200 -- B200-native quantized forward path behind the QuantTorch surface.

    import pytorch_quantize_impls_b200 as QuantTorch
    QuantTorch.layers.LinearBin(...), QuantTorch.functions.BinaryConnect(), QuantTorch.BinaryNet.LinearBin ...

Unlike the reference's package init (QuantTorch/__init__.py:1-7) nothing here imports the training / optuna
tooling; only the hot path (functions, layers, facades) exists.
"""
from . import _lib  # noqa: F401  (defines the loud failure when libqtb200.so is missing)
from . import functions, layers  # noqa: F401
from . import BinaryNet, DorefaNet, LogLinNet, TernerNet, XnorNet  # noqa: F401
from ._engine import (code_only_activations, set_backend, set_banded_head, set_first_layer_implicit,  # noqa: F401
                      set_first_layer_windows, set_overlap_head,
                      set_first_layer_planes, set_fp4, set_grad_backend, set_implicit_conv, set_wfold, set_xnor_mode)
from ._ops import device_caps, set_strict  # noqa: F401
from .fusion import (FlattenCodes, FusedActLayer, FusedBasicBlock, FusedBNActQuant, FusedConvPool, FusedLayerBN,  # noqa: F401
                     FusedLayerPoolQuant, FusedLayerQuant, OperandPrefetch, fuse_inference, prefetch_operands)
from .device import device  # noqa: F401
from .checkpoint import load_packed, packed_state, save_packed  # noqa: F401
from . import convertor  # noqa: F401

__version__ = '0.1'
