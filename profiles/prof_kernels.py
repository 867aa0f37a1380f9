"""Workload for the ncu captures in this directory (one GPU):
    ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|act_quant|weight_expand|simt_gemm' -c 16 \
        -o gpurun_out/prof python profiles/prof_kernels.py
Launch order (north-star shapes, M 8192, K = N = 4096 unless stated; one warm-up pass of everything first, not captured
because of -s):
  1. the BASELINE configs[1] step as bench.py runs it: XnorNet MLP 4096-4096-4096-1000, fuse_inference + code-only
     (one-pass XNOR quantizer, 3 x [weight expansion, tcgen05 cta_group::2 kind::f16 GEMM with the requant epilogue]);
  2. BinaryConnect -> LinearBin in drop-in mode (SIGN quantizer writing fp32 + e2m1 codes, expansion, cta_group::2 kind::mxf4);
  3. DorefaQuant(4) -> LinearDorefa(4) (DOREFA quantizer, expansion, cta_group::2 kind::i8);
  4. the CUDA-core XNOR+popcount kernel on the same LinearBin;
  5. DoReFa-4 conv 256->256 3x3 on 256 x 28 x 28 (channels-last quantizer, TMA-im2col implicit GEMM, cta_group::1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pytorch_quantize_impls_b200 as Q
from pytorch_quantize_impls_b200 import nets

dev = "cuda"
M, K, N = 8192, 4096, 4096
torch.manual_seed(0)
x = torch.randn(M, K, device=dev)
xu = torch.rand(M, K, device=dev)
xi = torch.rand(256, 256, 28, 28, device=dev)
with torch.no_grad():
    mlp = Q.fuse_inference(nets.xnor_mlp().to(dev).eval())
    lay = Q.layers.LinearBin(K, N).to(dev).eval()
    act = Q.functions.BinaryConnect()
    ld = Q.layers.LinearDorefa(K, N, bit_width=4).to(dev).eval()
    conv = Q.layers.DorefaConv2d(256, 256, 3, padding=1, bit_width=4).to(dev).eval()

    def everything():
        with Q.code_only_activations():
            mlp(x)
        lay(act(x))
        ld(Q.functions.DorefaQuant(xu, 4))
        Q.set_backend(popcount=True)
        lay(act(x))
        Q.set_backend(popcount=False)
        conv(Q.functions.DorefaQuant(xi, 4))

    everything()          # warm-up (skipped by ncu -s)
    torch.cuda.synchronize()
    everything()
torch.cuda.synchronize()
