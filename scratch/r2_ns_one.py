"""One north-star layer (BinaryConnect -> LinearBin 4096x4096, batch 8192, code-only) a few times: the workload of the ncu capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pytorch_quantize_impls_b200 as Q

dev = torch.device("cuda")
g = torch.Generator().manual_seed(99)
xs = [torch.randn(8192, 4096, generator=g).to(dev) for _ in range(2)]
lay = Q.layers.LinearBin(4096, 4096).to(dev).eval()
act = Q.functions.BinaryConnect()
with torch.no_grad(), Q.code_only_activations():
    for i in range(4):
        y = lay(act(xs[i % 2]))
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
