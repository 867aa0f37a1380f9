"""BinaryNet facade -- BinaryNet (Courbariaux et al.): sign activations / weights.

One import gives a model file every op and layer of the family, as `QuantTorch/BinaryNet.py:1-2` does for the reference
(`from QuantTorch.BinaryNet import LinearX, ...`).  The names are listed explicitly (no star import), so that what a drop-in
user can rely on is visible here and checked by tests/test_cabi_and_surface.py.
"""
from .functions.binary_connect import (  # noqa: F401
    BinaryConnect, BinaryConnectDeterministic, BinaryConnectStochastic, BinaryConv2d, BinaryDense,
    TaggingFunction, front, safeSign, ste_clip,
)
from .layers.binary_layers import (  # noqa: F401
    BinConv2d, LinearBin, QuantLayerMixin, check_convert,
)

__all__ = sorted(n for n in dir() if not n.startswith("_"))
