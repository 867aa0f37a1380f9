"""Facade, QuantTorch/BinaryNet.py:1-2."""
from .functions.binary_connect import *  # noqa
from .layers.binary_layers import *  # noqa
