"""Facade, QuantTorch/TernerNet.py:1-2."""
from .functions.terner_connect import *  # noqa
from .layers.terner_layers import *  # noqa
