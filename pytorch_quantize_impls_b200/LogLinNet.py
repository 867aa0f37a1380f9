"""LogLinNet facade -- Miyashita et al. lin / log fixed-point quantization.

One import gives a model file every op and layer of the family, as `QuantTorch/LogLinNet.py:1-2` does for the reference
(`from QuantTorch.LogLinNet import LinearX, ...`).  The names are listed explicitly (no star import), so that what a drop-in
user can rely on is visible here and checked by tests/test_cabi_and_surface.py.
"""
from .functions.log_lin_connect import (  # noqa: F401
    LinQuant, LogQuant, Quant, TaggingFunction, front, nnQuant,
)
from .layers.log_lin_layers import (  # noqa: F401
    LinearQuant, QuantConv2d, QuantLayerMixin, check_convert,
)

__all__ = sorted(n for n in dir() if not n.startswith("_"))
