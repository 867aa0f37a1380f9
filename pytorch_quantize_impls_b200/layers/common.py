"""Shared machinery of the quantized layers.

The reference repeats the same train()/forward pattern in every layer class
(binary_layers.py:30-46, terner_layers.py:30-51, dorefa_layers.py:29-45, log_lin_layers.py:22-42);
here it lives once, parameterised by two hooks:
    _weight_op(w)   differentiable fake-quant of the fp32 weights (the reference's bin_op / ter_op / weight_op)
    _make_pack(w)   k-bit HBM pack of the fp32 master weights
"""
import torch

from .. import _engine as eng
from .. import _ops as ops


class QLayer():
    """Marker base class, QuantTorch/layers/common.py:1-9."""

    def get_quant_weight(self):
        raise NotImplementedError

    def set_quant_weight(self):
        raise NotImplementedError

    def restore_weight(self):
        raise NotImplementedError


class _EvalState:
    __slots__ = ("pack", "version", "ptr")


class _Contraction(torch.autograd.Function):
    """Forward = low-bit kernels; backward = the reference's autograd semantics (F.linear / F.conv2d of the
    fake-quantized weight, STE through the weight op), evaluated with torch ops."""

    @staticmethod
    def forward(ctx, input, weight, bias, layer):
        ctx.layer = layer
        # the mode and (for stochastic layers) the drawn weights of THIS forward are what backward must see: the reference's
        # autograd graph keeps the sampled tensor, it does not draw again
        ctx.was_training = layer.training
        pack, wq_sample = layer._pack_for_forward()
        ctx.wq_sample = wq_sample
        ctx.save_for_backward(input, weight, bias)
        return layer._run_kernels(input, pack)

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, bias = ctx.saved_tensors
        layer = ctx.layer
        gi = gw = gb = None
        with torch.enable_grad():
            w = weight.detach().requires_grad_(True)
            # stochastic ops: the graph only serves the STE (a function of w alone); the VALUES come from the forward's sample
            wq = layer._weight_op(w) if ctx.was_training else w
        wqd = ctx.wq_sample if ctx.wq_sample is not None else wq.detach()
        if layer._is_conv:
            kw = dict(stride=layer.stride, padding=layer.padding, dilation=layer.dilation, groups=layer.groups)
            # both gradient contractions on the bf16 hi/lo tensor-core route (engine.grad_*_conv2d)
            if ctx.needs_input_grad[0]:
                gi = eng.grad_input_conv2d(input.shape, wqd, grad_output, **kw)
            if ctx.needs_input_grad[1]:
                gwq = eng.grad_weight_conv2d(input, weight.shape, grad_output, **kw)
            if bias is not None and ctx.needs_input_grad[2]:
                gb = grad_output.sum((0, 2, 3))
        else:
            # gradient contractions on the bf16 tensor-core route (engine.grad_*: hi/lo planes, ~2^-16 relative error)
            g2 = grad_output.reshape(-1, grad_output.shape[-1])
            if ctx.needs_input_grad[0]:
                gi = eng.grad_input_linear(g2, wqd).reshape(input.shape)
            if ctx.needs_input_grad[1]:
                gwq = eng.grad_weight_linear(g2, input.reshape(-1, input.shape[-1]))
            if bias is not None and ctx.needs_input_grad[2]:
                gb = g2.sum(0)
        if ctx.needs_input_grad[1]:
            if wq is w:
                gw = gwq
            else:
                gw, = torch.autograd.grad(wq, w, gwq)
        return gi, gw, gb, None


class QuantLayerMixin(QLayer):
    _is_conv = False
    _eval_state = None
    _packed_only = None       # (WeightPack, weight shape) installed by checkpoint.load_packed: no fp32 master weights
    _in_perm = None           # (C, H, W): this Linear reads a channels-last flatten (fusion.FlattenCodes) of a [C, H, W] activation

    def _w2d(self, w):
        """[out, K] matrix the packers read: conv weights with the channel fastest (the K order of the im2col kernels); a Linear
        fed by channels-last flattened codes gets its columns permuted from (c, h, w) to (h, w, c) order."""
        if self._in_perm is not None and w.dim() == 2:
            C_, H_, W_ = self._in_perm
            return w.reshape(w.shape[0], C_, H_, W_).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
        return ops.conv_weight_2d(w)

    def _set_in_perm(self, perm):
        if self._packed_only is not None:
            raise RuntimeError("packed-only layers cannot be re-ordered")
        self._in_perm = perm
        self._eval_state = None

    def _wshape(self):
        return self._packed_only[1] if self._packed_only is not None else tuple(self.weight.shape)

    def _install_packed(self, pack, weight_shape, drop_master=True):
        """Make the layer packed-only (inference): the kernels read `pack`; the fp32 weights are released."""
        self._packed_only = (pack, tuple(weight_shape))
        self._eval_state = None
        self.training = False
        if drop_master:
            if hasattr(self.weight, "org"):
                del self.weight.org
            self.weight = torch.nn.Parameter(torch.empty(0, device=self.weight.device), requires_grad=False)

    # ---- hooks -------------------------------------------------------------------------
    def _weight_op(self, w):
        raise NotImplementedError

    def _make_pack(self, w):
        raise NotImplementedError

    def _weight_op_host(self, w):
        """The weight op as plain torch arithmetic, for the train(False) swap of a layer whose weights are not on a CUDA
        device yet (`model.eval(); model.cuda()`): module state management, not a contraction path -- forward on a CPU
        tensor still raises."""
        raise NotImplementedError

    def _is_stochastic(self):
        return not getattr(self, "deterministic", True)

    def _make_pack_of_sample(self, wq):
        """Pack of an already drawn stochastic sample (values in the quantizer's own output set)."""
        raise NotImplementedError

    def _pack_for_forward(self):
        """(WeightPack, drawn weights or None) for one forward.  Stochastic layers draw ONCE here: the pack the kernels
        contract with and the tensor backward multiplies the output gradient by are the same sample."""
        if self._packed_only is None and self.training and self._is_stochastic():
            with torch.no_grad():
                wq = self._weight_op(self.weight.detach())
            return self._make_pack_of_sample(wq), wq
        return self._current_pack(), None

    # ---- weight-swap train()/eval(), e.g. binary_layers.py:30-40 --------------------------
    def train(self, mode=True):
        if self._packed_only is not None:
            if mode:
                raise RuntimeError("this layer holds packed k-bit weights only (checkpoint.load_packed): load fp32 master "
                                   "weights before training")
            return self
        if self.training == mode:
            return self
        if mode:
            self.weight.data.copy_(self.weight.org.data)
            self._eval_state = None
            self.training = True
            return self
        # eval: nothing of the layer's state changes until quantisation and packing have succeeded
        master = self.weight.data.clone()
        st = None
        with torch.no_grad():
            if self.weight.is_cuda:
                st = _EvalState()
                if self._is_stochastic():
                    wq = self._weight_op(self.weight).detach()        # ONE draw: stored weights and pack agree
                    st.pack = self._make_pack_of_sample(wq)
                else:
                    st.pack = self._make_pack(self.weight)            # packed once, from the fp32 master weights
                    wq = self._weight_op(self.weight).detach()
            else:
                # weights still on the host: swap with torch arithmetic, pack on the first CUDA forward (_current_pack
                # re-packs from `weight.org` / the stored quantized values)
                wq = self._weight_op_host(self.weight.detach())
        if not hasattr(self.weight, 'org'):
            self.weight.org = master
        else:
            self.weight.org.data = master
        self.weight.data.copy_(wq)
        if st is not None:
            st.version, st.ptr = self.weight._version, self.weight.data_ptr()
        self._eval_state = st
        self.training = False
        return self

    def _current_pack(self):
        if self._packed_only is not None:
            return self._packed_only[0]
        if self.training:
            if self._is_stochastic():
                return self._pack_for_forward()[0]
            return self._make_pack(self.weight)               # the reference re-quantizes W on every call
        st = self._eval_state
        if st is not None and st.version == self.weight._version and st.ptr == self.weight.data_ptr():
            return st.pack
        # eval mode, but the cached pack no longer describes `weight` (module moved with .to(), deep-copied, state_dict
        # loaded in eval mode, ...).  If the fp32 master copy made by train(False) is still around and the stored weights
        # are its quantisation, re-pack from it; if the stored values are themselves a fixed point of the quantizer
        # (+-1 / {-1,0,1}), pack those; otherwise contract with the stored values as they are (what the reference does
        # in eval mode) through the real-valued weight operand.
        st = _EvalState()
        st.pack = None
        w = self.weight.detach()
        org = getattr(self.weight, "org", None)
        with torch.no_grad():
            if self._is_stochastic():
                # the stored values ARE the sample drawn at train(False); never draw again in eval mode
                st.pack = self._make_pack_of_sample(w)
            else:
                if org is not None and org.shape == w.shape:
                    org = org.to(w.device)
                    if torch.equal(self._weight_op(org), w):
                        st.pack = self._make_pack(org)
                if st.pack is None and torch.equal(self._weight_op(w), w):
                    st.pack = self._make_pack(w)
        if st.pack is None:
            st.pack = ops.pack_real_weight(self._w2d(w))
        st.version, st.ptr = self.weight._version, self.weight.data_ptr()
        self._eval_state = st
        return st.pack

    def _run_kernels(self, input, pack=None):
        if pack is None:
            pack = self._pack_for_forward()[0]
        if self._is_conv:
            return eng.conv2d(input, pack, self.bias, self._wshape(), self.stride, self.padding,
                              self.dilation, self.groups)
        return eng.linear(input, pack, self.bias)

    def _forward_requant(self, input, spec, **conv_kw):
        """Inference chain: contraction with the next activation quantizer fused into the epilogue (fusion.FusedLayerQuant)."""
        eng.tagged_input_device(input)
        pack = self._current_pack()
        if self._is_conv:
            return eng.conv2d(input, pack, self.bias, self._wshape(), self.stride, self.padding,
                              self.dilation, self.groups, requant=spec, **conv_kw)
        return eng.linear(input, pack, self.bias, requant=spec)

    def _forward_affine(self, input, spec, **conv_kw):
        """Inference: contraction with a per-output-channel affine (folded eval-mode BatchNorm) in the epilogue."""
        eng.tagged_input_device(input)
        pack = self._current_pack()
        if self._is_conv:
            return eng.conv2d(input, pack, self.bias, self._wshape(), self.stride, self.padding,
                              self.dilation, self.groups, affine=spec, **conv_kw)
        return eng.linear(input, pack, self.bias, affine=spec)

    def forward(self, input):
        eng.tagged_input_device(input)
        if self._packed_only is not None:
            return self._run_kernels(input)
        needs_grad = torch.is_grad_enabled() and (
            input.requires_grad or self.weight.requires_grad or (self.bias is not None and self.bias.requires_grad))
        if needs_grad:
            return _Contraction.apply(input, self.weight, self.bias, self)
        return self._run_kernels(input)


def check_convert(other, cls, name):
    if not isinstance(other, cls):
        raise TypeError("Expected a {} ! Receive:  {}".format(name, other.__class__))
