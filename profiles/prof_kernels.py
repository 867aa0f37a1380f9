"""Workload for the ncu captures in this directory (one GPU):
    ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|act_quant|simt_gemm' -c 12 -o gpurun_out/prof python profiles/prof_kernels.py
Launch order: BinaryConnect quantizer, LinearBin (kind::i8 tcgen05), XNOR quantizer, LinearXNOR (kind::f16 tcgen05),
XNOR+popcount CUDA-core kernel, DoReFa-4 conv (TMA-im2col implicit GEMM) -- north-star shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pytorch_quantize_impls_b200 as Q

dev = "cuda"
M, K, N = 8192, 4096, 4096
torch.manual_seed(0)
x = torch.randn(M, K, device=dev)
with torch.no_grad():
    lay = Q.layers.LinearBin(K, N).to(dev).eval()
    act = Q.functions.BinaryConnect()
    for _ in range(2):
        y = lay(act(x))
    lx = Q.layers.LinearXNOR(K, N).to(dev).eval()
    for _ in range(2):
        y = lx(Q.functions.QuantXnor(x, 1))
    Q.set_backend(popcount=True)
    y = lay(act(x))
    Q.set_backend(popcount=False)
    conv = Q.layers.DorefaConv2d(256, 256, 3, padding=1, bit_width=4).to(dev).eval()
    xi = torch.rand(256, 256, 28, 28, device=dev)
    for _ in range(2):
        y = conv(Q.functions.DorefaQuant(xi, 4))
torch.cuda.synchronize()
