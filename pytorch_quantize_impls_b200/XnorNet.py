"""Facade, QuantTorch/XnorNet.py:1-2."""
from .functions.xnor_connect import *  # noqa
from .layers.xnor_layers import *  # noqa
