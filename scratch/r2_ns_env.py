"""North-star contraction (gemm_f4 on sign codes, 8192 x 4096 x 4096) under the epilogue switches of the moment (env)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import pytorch_quantize_impls_b200 as Q

dev = torch.device("cuda")
g = torch.Generator().manual_seed(99)
x = torch.randn(8192, 4096, generator=g).to(dev)
lay = Q.layers.LinearBin(4096, 4096).to(dev).eval()
act = Q.functions.BinaryConnect()
with torch.no_grad():
    xq = act(x)
    ms = bench.time_fn(torch, lambda: lay(xq), iters=40, graph=True)
print({k: v for k, v in os.environ.items() if k.startswith("QTB200")}, "contraction us", round(ms * 1e3, 2))
