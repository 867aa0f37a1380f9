"""XNOR-Net layers -- surface of QuantTorch/layers/xnor_layers.py.

The reference's LinearXNOR.train(False) raises NameError (`dim` instead of `self.dim`, xnor_layers.py:30); here
the swap works (W <- mean(|W|, DIM=0) * sign(W), which is what XNORDense computes) and, as in the reference,
forward applies the XNOR op in both modes."""
import torch

from .. import _engine as eng
from .. import _ops as ops
from ..functions import xnor_connect
from .common import QuantLayerMixin, _EvalState, check_convert


class _XnorMixin(QuantLayerMixin):
    def forward(self, input):
        eng.tagged_input_device(input)
        if self._packed_only is not None:
            return self._run_kernels(input)
        pack = None if self.training else self._current_pack()
        op = self.conv_op if self._is_conv else self.lin_op
        return op.apply(input, self.weight, self.bias, pack)

    def _current_pack(self):
        if self._packed_only is not None:
            return self._packed_only[0]
        st = self._eval_state
        if (not self.training and st is not None and st.version == self.weight._version
                and st.ptr == self.weight.data_ptr()):
            return st.pack
        pack = self._make_pack(self.weight)
        if not self.training:
            # eval mode with a stale / missing cache (module moved with .to(), eval() before .cuda(), state_dict loaded):
            # forward re-applies the op to the stored weights in both modes (reference behaviour), so pack those once
            st = _EvalState()
            st.pack, st.version, st.ptr = pack, self.weight._version, self.weight.data_ptr()
            self._eval_state = st
        return pack

    def train(self, mode=True):
        if self._packed_only is not None:
            return QuantLayerMixin.train(self, mode)
        if self.training == mode:
            return self
        if mode:
            self.weight.data.copy_(self.weight.org.data)
            self._eval_state = None
            self.training = True
            return self
        master = self.weight.data.clone()
        with torch.no_grad():
            wq = self._weight_op(self.weight).detach()          # torch arithmetic: works wherever the weights live
        if not hasattr(self.weight, 'org'):
            self.weight.org = master
        else:
            self.weight.org.data = master
        self.weight.data.copy_(wq)
        self._eval_state = None                                  # packed on the first eval-mode forward (_current_pack)
        self.training = False
        return self


class LinearXNOR(_XnorMixin, torch.nn.Linear):
    """y = x . (sign(W) * alpha)^T + b, alpha[k] = mean(|W|, 0): weights are 2 bit planes + alpha (xnor_layers.py:8-34)."""

    @staticmethod
    def convert(other, dim=[0, 1]):
        check_convert(other, torch.nn.Linear, "torch.nn.Linear")
        return LinearXNOR(other.in_features, other.out_features, False if other.bias is None else True, dim=dim)

    def __init__(self, in_features, out_features, bias=True, dim=[0, 1]):
        super().__init__(in_features, out_features, bias=bias)
        self.lin_op = xnor_connect.XNORDense(dim=dim)
        self.dim = dim

    def _weight_op(self, w):
        return torch.mean(torch.abs(w), xnor_connect.DIM, keepdim=True) * torch.sign(w)

    def _make_pack(self, w):
        return xnor_connect.xnor_pack(w)


class XNORConv2d(_XnorMixin, torch.nn.Conv2d):
    """conv2d(x, sign(W) * mean(|W|, dim)) + b (xnor_layers.py:36-69); `quant_input` is ignored as in the
    reference (xnor_layers.py:49 passes False)."""
    _is_conv = True

    @staticmethod
    def convert(other, dim=[0, 1], quant_input=False):
        check_convert(other, torch.nn.Conv2d, "torch.nn.Conv2d")
        return XNORConv2d(other.in_channels, other.out_channels, other.kernel_size, stride=other.stride,
                          padding=other.padding, dilation=other.dilation, groups=other.groups,
                          bias=False if other.bias is None else True, dim=dim, quant_input=quant_input)

    def __init__(self, *kargs, dim=[0, 1], quant_input=False, **kwargs):
        torch.nn.Conv2d.__init__(self, *kargs, **kwargs)
        self.dim = dim
        self.conv_op = xnor_connect.XNORConv2d(dim, False, self.stride, self.padding, self.dilation, self.groups)

    def _weight_op(self, w):
        return torch.mean(torch.abs(w), self.dim, keepdim=True) * torch.sign(w)

    def _make_pack(self, w):
        return xnor_connect.xnor_conv_pack(w, self.dim)

    def clamp(self):
        pass
