"""Same export list as QuantTorch/layers/__init__.py:1-28, minus the Elastic/WQR layers (out of scope)."""
from .binary_layers import BinConv2d, LinearBin
from .common import QLayer
from .dorefa_layers import DorefaConv2d, LinearDorefa
from .log_lin_layers import LinearQuant, QuantConv2d
from .terner_layers import LinearTer, TerConv2d
from .xnor_layers import LinearXNOR, XNORConv2d
