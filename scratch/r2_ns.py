"""North-star layer timing probe (BinaryConnect -> LinearBin 4096x4096, batch 8192): plain code-only pair vs the banded pipeline."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import pytorch_quantize_impls_b200 as Q
from pytorch_quantize_impls_b200 import _engine as eng

dev = torch.device("cuda")
M = K = N = None
M, K, N = 8192, 4096, 4096
g = torch.Generator().manual_seed(99)
xs = [torch.randn(M, K, generator=g).to(dev) for _ in range(3)]
lay = Q.layers.LinearBin(K, N).to(dev)
lay.bias.data.uniform_(-1, 1)
lay.eval()
act = Q.functions.BinaryConnect()
pair = Q.fuse_inference(torch.nn.Sequential(act, lay))
i = [0]
bytes_module = 4.0 * M * K + N * K / 8 + 4 * N + 4.0 * M * N
out = {}


def rec(name, ms):
    out[name] = {"us": round(ms * 1e3, 2), "hbm_frac": round(bytes_module / ms / 1e6 / 6454.0, 4)}


with torch.no_grad():
    def f_plain():
        i[0] += 1
        with Q.code_only_activations():
            return lay(act(xs[i[0] % 3]))
    rec("code_only_plain", bench.time_fn(torch, f_plain, iters=40, graph=True))
    xq = act(xs[0])
    rec("contraction_only", bench.time_fn(torch, lambda: lay(xq), iters=40, graph=True))
    with Q.code_only_activations():
        xqc = act(xs[0])
        rec("gemm_f4_on_codes", bench.time_fn(torch, lambda: lay(xqc), iters=40, graph=True))

    def f_q():
        i[0] += 1
        with Q.code_only_activations():
            return act(xs[i[0] % 3])
    rec("quantizer_code_only", bench.time_fn(torch, f_q, iters=40, graph=True))
    def f_pair():
        i[0] += 1
        with Q.code_only_activations():
            return pair(xs[i[0] % 3])

    Q.set_overlap_head(False)
    i[0] = 0
    ref = f_pair().clone()                                   # plain pair: one kernel after the other
    Q.set_overlap_head(True)                                 # quantizer beside the contraction (progress counters)
    i[0] = 0
    y = f_pair()
    out["overlapped_equals_plain"] = bool(torch.equal(y, ref))
    rec("overlapped", bench.time_fn(torch, f_pair, iters=40, graph=True))
    Q.set_overlap_head(False)
    Q.set_banded_head(True)
    for nb in (2, 4, 8):
        orig = eng.linear_banded

        def banded(x, quantize, pack, bias, nbands=4, affine=None, _nb=nb):
            return orig(x, quantize, pack, bias, nbands=_nb, affine=affine)
        eng.linear_banded = banded
        f_pair()
        rec("banded_%d" % nb, bench.time_fn(torch, f_pair, iters=40, graph=True))
        eng.linear_banded = orig
    Q.set_banded_head(False)
    Q.set_overlap_head(True)
print(json.dumps(out, indent=1))
