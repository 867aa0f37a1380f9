"""TernerNet facade -- ternary weight networks: {-1, 0, +1} weights.

One import gives a model file every op and layer of the family, as `QuantTorch/TernerNet.py:1-2` does for the reference
(`from QuantTorch.TernerNet import LinearX, ...`).  The names are listed explicitly (no star import), so that what a drop-in
user can rely on is visible here and checked by tests/test_cabi_and_surface.py.
"""
from .functions.terner_connect import (  # noqa: F401
    TaggingFunction, TernaryConnect, TernaryConnectDeterministic, TernaryConnectStochastic, TernaryConv2d,
    TernaryDense, front, safeSign, ste_clip,
)
from .layers.terner_layers import (  # noqa: F401
    LinearTer, QuantLayerMixin, TerConv2d, check_convert, sqrt,
)

__all__ = sorted(n for n in dir() if not n.startswith("_"))
