"""Isolated timings of the small kernels of the step (L2 flushed between launches by a 256 MB memset)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_quantize_impls_b200 as Q
from pytorch_quantize_impls_b200 import _lib as L, _ops as ops, _engine as eng

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10, do_flush=True):
    if not do_flush:      # back-to-back launches (callers rotate their inputs), one event pair
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3 * n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / (3 * n) * 1e3
    ts = []
    for i in range(n + 2):
        if do_flush:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


M, K, N = 8192, 4096, 4096
x = torch.randn(M, K, device=dev)
xu = torch.rand(M, K, device=dev)
xs = [torch.randn(M, K, device=dev) for _ in range(3)]
_i = [0]


def rot():
    _i[0] += 1
    return xs[_i[0] % 3]

w = torch.randn(N, K, device=dev) * 0.02
px = ops.pack_weight(w, "xnor")
ps = ops.pack_weight(w, "sign")
pd = ops.pack_weight(w, "dorefa", 4)


def gtime(mk, n_in_graph=6, reps=5):
    """mk(i) launches the kernel on rotating input i; 6 launches captured in one CUDA graph, replayed: no host gaps."""
    for i in range(3):
        mk(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.graph(g):
        for i in range(n_in_graph):
            keep.append(mk(i))
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / (reps * n_in_graph) * 1e3


print("-- CUDA-graph replays, rotating 3 inputs (QTB200_QCHUNK=%s)" % os.environ.get("QTB200_QCHUNK"))
print("quant xnor code-only   %.1f us" % gtime(lambda i: ops.quant_act(xs[i % 3], L.Q_XNOR_ROW, want_y=False, codes_kind=L.CODES_F16, want_row_scale=True, kind="xnor")))
print("quant xnor drop-in     %.1f us" % gtime(lambda i: ops.quant_act(xs[i % 3], L.Q_XNOR_ROW, want_y=True, codes_kind=L.CODES_F16, want_row_scale=True, kind="xnor")))
print("quant sign code-only   %.1f us" % gtime(lambda i: ops.quant_act(xs[i % 3], L.Q_SIGN, want_y=False, codes_kind=L.CODES_F4, kind="sign")))
print("quant sign drop-in     %.1f us" % gtime(lambda i: ops.quant_act(xs[i % 3], L.Q_SIGN, want_y=True, codes_kind=L.CODES_F4, want_bits=True, kind="sign")))
print("quant dorefa4 code-only %.1f us" % gtime(lambda i: ops.quant_act(xs[i % 3].abs(), L.Q_DOREFA, bit_width=4, want_y=False, codes_kind=L.CODES_I8, want_row_sum=True, kind="dorefa")))
print("expand xnor->fp16      %.1f us" % gtime(lambda i: ops.expand_weight(px, L.CODES_F16)))
print("expand sign->f4        %.1f us" % gtime(lambda i: ops.expand_weight(ps, L.CODES_F4)))
print("expand sign->i8        %.1f us" % gtime(lambda i: ops.expand_weight(ps, L.CODES_I8)))
print("expand dorefa4->i8     %.1f us" % gtime(lambda i: ops.expand_weight(pd, L.CODES_I8)))
