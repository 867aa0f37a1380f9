"""DorefaNet facade -- DoReFa-Net (Zhou et al.): k-bit weights and activations.

One import gives a model file every op and layer of the family, as `QuantTorch/DorefaNet.py:1-2` does for the reference
(`from QuantTorch.DorefaNet import LinearX, ...`).  The names are listed explicitly (no star import), so that what a drop-in
user can rely on is visible here and checked by tests/test_cabi_and_surface.py.
"""
from .functions.dorefa_connect import (  # noqa: F401
    DorefaQuant, QuantConv2d, QuantDense, TaggingFunction, dorefa_pack, front, nnDorefaQuant, nnQuantWeight,
    safeSign,
)
from .layers.dorefa_layers import (  # noqa: F401
    DorefaConv2d, LinearDorefa, QuantLayerMixin, check_convert,
)

__all__ = sorted(n for n in dir() if not n.startswith("_"))
