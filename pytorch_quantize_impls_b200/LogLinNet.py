"""Facade, QuantTorch/LogLinNet.py:1-2."""
from .functions.log_lin_connect import *  # noqa
from .layers.log_lin_layers import *  # noqa
