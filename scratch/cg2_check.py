"""cta_group::2 GEMM kernels against the cta_group::1 kernels (bit-exact for the integer kinds) + timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_quantize_impls_b200 as Q
from pytorch_quantize_impls_b200 import _lib as L, _ops as ops

dev = torch.device("cuda")


def run_i8(M, N, K, cg):
    L.set_option("cta_group", cg)
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randint(-3, 4, (M, K), generator=g, dtype=torch.int8).to(dev)
    w = torch.randint(-3, 4, (N, K), generator=g, dtype=torch.int8).to(dev)
    out = torch.empty(M, N, device=dev)
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    bias = torch.randn(N, generator=g).to(dev)
    epi = ops.make_epi(out, ldo=N, bias=bias, acc_out=acc)
    ops.gemm_i8(a, True, K, w, True, K, M, N, K, epi, L.BACKEND_TCGEN05)
    torch.cuda.synchronize()
    return acc, out, (a, w)


def run_f16(M, N, K, cg):
    L.set_option("cta_group", cg)
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).half().to(dev)
    w = torch.randn(N, K, generator=g).half().to(dev)
    out = torch.empty(M, N, device=dev)
    epi = ops.make_epi(out, ldo=N)
    ops.gemm_f16(a, K, 0, w, K, 0, [(0, 0)], M, N, K, epi, L.BACKEND_TCGEN05, fmt=L.FMT_FP16)
    torch.cuda.synchronize()
    return out, (a, w)


ok = True
for (M, N, K) in [(256, 256, 128), (1024, 512, 256), (1000, 1000, 1008), (300, 272, 4096), (4096, 4096, 4096), (129, 600, 64)]:
    acc1, out1, (a, w) = run_i8(M, N, K, 1)
    acc2, out2, _ = run_i8(M, N, K, 2)
    ref = (a.float() @ w.float().t()).to(torch.int32)
    e1 = torch.equal(acc1, acc2) and torch.equal(out1, out2)
    e2 = torch.equal(acc2, ref)
    print("i8", M, N, K, "cg2==cg1", e1, "cg2==ref", e2, flush=True)
    ok &= e1 and e2
    o1, (a, w) = run_f16(M, N, K, 1)
    o2, _ = run_f16(M, N, K, 2)
    e = torch.equal(o1, o2)
    r = float((o2 - a.float() @ w.float().t()).abs().max() / (a.float() @ w.float().t()).abs().max())
    print("f16", M, N, K, "cg2==cg1", e, "rel vs torch", r, flush=True)
    ok &= r < 1e-3


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


M, N, K = 8192, 4096, 4096
a8 = torch.randint(-1, 2, (M, K), dtype=torch.int8).to(dev)
w8 = torch.randint(-1, 2, (N, K), dtype=torch.int8).to(dev)
ah = torch.randn(M, K).half().to(dev)
wh = torch.randn(N, K).half().to(dev)
a4 = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8).to(dev) & 0xAA    # nibbles 0xA / 0x2 / 0x8 / 0: +-1 and 0 codes
w4 = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8).to(dev) & 0xAA
out = torch.empty(M, N, device=dev)
outs = {}
for cg in (1, 2):
    L.set_option("cta_group", cg)
    epi = ops.make_epi(out, ldo=N)
    t_i8 = timeit(lambda: ops.gemm_i8(a8, True, K, w8, True, K, M, N, K, epi, L.BACKEND_TCGEN05))
    t_f16 = timeit(lambda: ops.gemm_f16(ah, K, 0, wh, K, 0, [(0, 0)], M, N, K, epi, L.BACKEND_TCGEN05, fmt=L.FMT_FP16))
    t_f4 = timeit(lambda: ops.gemm_f4(a4, K, w4, K, M, N, K, epi))
    outs[cg] = out.clone()
    ops_ = 2.0 * M * N * K
    print("cta_group %d: i8 %.4f ms (%.0f TOPS)  f16 %.4f ms (%.0f TF)  f4 %.4f ms (%.0f TOPS)" % (
        cg, t_i8, ops_ / t_i8 / 1e9, t_f16, ops_ / t_f16 / 1e9, t_f4, ops_ / t_f4 / 1e9), flush=True)
print("f4 cg2==cg1", torch.equal(outs[1], outs[2]))
ok &= torch.equal(outs[1], outs[2])
L.set_option("cta_group", 0)
print("ALL OK" if ok else "MISMATCH")
