"""XnorNet facade -- XNOR-Net (Rastegari et al.): sign * mean activations, alpha-scaled sign weights.

One import gives a model file every op and layer of the family, as `QuantTorch/XnorNet.py:1-2` does for the reference
(`from QuantTorch.XnorNet import LinearX, ...`).  The names are listed explicitly (no star import), so that what a drop-in
user can rely on is visible here and checked by tests/test_cabi_and_surface.py.
"""
from .functions.xnor_connect import (  # noqa: F401
    DIM, QuantXnor, TaggingFunction, XNORConv2d, XNORDense, front, nnQuantXnor, xnor_conv_pack, xnor_pack,
)
from .layers.xnor_layers import (  # noqa: F401
    LinearXNOR, QuantLayerMixin, XNORConv2d, check_convert,
)

__all__ = sorted(n for n in dir() if not n.startswith("_"))
