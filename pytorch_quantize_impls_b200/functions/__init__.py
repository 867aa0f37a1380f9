"""Same export list as QuantTorch/functions/__init__.py:1-51, minus the Elastic/WQR regularisers
(out of scope: training-time penalties whose forward is a plain fp32 GEMM)."""
from .binary_connect import (BinaryConnect, BinaryConnectDeterministic, BinaryConnectStochastic, BinaryConv2d,
                             BinaryDense)
from .common import front, front2, safeSign
from .dorefa_connect import DorefaQuant, QuantConv2d, QuantDense, nnDorefaQuant, nnQuantWeight
from .log_lin_connect import LinQuant, LogQuant, Quant, nnQuant
from .terner_connect import (TernaryConnect, TernaryConnectDeterministic, TernaryConnectStochastic, TernaryConv2d,
                             TernaryDense)
from .xnor_connect import QuantXnor, XNORConv2d, XNORDense, nnQuantXnor
