"""mxf4 / f16 / i8 GEMM at the north-star shape with the epilogue switched off in stages (QTB200_EPI_DEBUG read at launch)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
dev = torch.device("cuda")
M, N, K = 8192, 4096, 4096
res = {}
a4 = (torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev) & 0xAA) | 0x22
w4 = (torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev) & 0xAA) | 0x22
ah = torch.randn(M, K, device=dev).half(); wh = torch.randn(N, K, device=dev).half()
ai = torch.randint(-3, 4, (M, K), dtype=torch.int8, device=dev); wi = torch.randint(-3, 4, (N, K), dtype=torch.int8, device=dev)
bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev)
for dbg in ("0", "3"):
    os.environ["QTB200_EPI_DEBUG"] = dbg
    res["f4_dbg" + dbg] = round(1e3 * bench.time_fn(torch, lambda: ops.gemm_f4(a4, K, w4, K, M, N, K, ops.make_epi(out, ldo=N, bias=bias)), iters=40, graph=True), 2)
    res["f16_dbg" + dbg] = round(1e3 * bench.time_fn(torch, lambda: ops.gemm_f16(ah, K, 0, wh, K, 0, [(0, 0)], M, N, K, ops.make_epi(out, ldo=N, bias=bias), L.BACKEND_TCGEN05, fmt=L.FMT_FP16), iters=20, graph=True), 2)
    res["i8_dbg" + dbg] = round(1e3 * bench.time_fn(torch, lambda: ops.gemm_i8(ai, True, K, wi, True, K, M, N, K, ops.make_epi(out, ldo=N, bias=bias), L.BACKEND_TCGEN05), iters=20, graph=True), 2)
os.environ["QTB200_EPI_DEBUG"] = "0"
print(json.dumps(res))
