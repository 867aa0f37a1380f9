"""qt_gemm_f4 timing scan (what bounds the e2m1 product at the north-star shape): K, N, tile and CTA-group variants, and the same
shape without an fp32 output (raw int32 accumulators only / requant codes only)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from pytorch_quantize_impls_b200 import _lib as L, _ops as ops

dev = torch.device("cuda")
res = {}


def codes(rows, K, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (rows, K // 2), generator=g, dtype=torch.uint8).to(dev) & 0xAA | 0x22      # nibbles in {2, A}: +-1


def run(tag, M, N, K, cg=0, out_kind="f32", f4bn=0):
    L.set_option("cta_group", cg)
    L.set_option("f4_tile_n", f4bn)
    a, w = codes(M, K, 1), codes(N, K, 2)
    bias = torch.randn(N).to(dev)
    outs = [torch.empty(M, N, device=dev) for _ in range(2)]
    i = [0]

    def f():
        i[0] += 1
        if out_kind == "f32":
            epi = ops.make_epi(outs[i[0] % 2], ldo=N, bias=bias)
        elif out_kind == "f32_nobias":
            epi = ops.make_epi(outs[i[0] % 2], ldo=N)
        else:
            rq = ops.RequantOut(L.Q_SIGN, 0, L.CODES_F4, M, N, dev)
            epi = ops.make_epi(None, ldo=N, bias=bias, requant=rq)
        ops.gemm_f4(a, K, w, K, M, N, K, epi)
    ms = bench.time_fn(torch, f, iters=40, graph=True)
    res[tag] = round(ms * 1e3, 2)
    print(tag, res[tag], flush=True)


run("base_8192x4096x4096", 8192, 4096, 4096)
run("K1024", 8192, 4096, 1024)
run("K2048", 8192, 4096, 2048)
run("K8192", 8192, 4096, 8192)
run("N4080", 8192, 4080, 4096)
run("N3840", 8192, 3840, 4096)
run("cg1_bn240", 8192, 4096, 4096, cg=1)
run("cg1_bn128", 8192, 4096, 4096, cg=1, f4bn=128)
run("nobias", 8192, 4096, 4096, out_kind="f32_nobias")
run("codes_out_only", 8192, 4096, 4096, out_kind="rq")
run("M4096", 4096, 4096, 4096)
L.set_option("cta_group", 0)
L.set_option("f4_tile_n", 0)
print(json.dumps(res))
