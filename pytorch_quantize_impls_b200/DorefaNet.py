"""Facade, QuantTorch/DorefaNet.py:1-2."""
from .functions.dorefa_connect import *  # noqa
from .layers.dorefa_layers import *  # noqa
