import sys, os, torch, numpy as np, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import quanttorch_oracle as O
import pytorch_quantize_impls_b200 as Q
from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
torch.manual_seed(0)
dev='cuda'
xu = torch.rand(256, 512)
ld = Q.layers.LinearDorefa(512, 384, bit_width=4)
w=ld.weight.data.clone(); b=ld.bias.data.clone()
refd = O.linear_dorefa(O.dorefa_quantize(xu, 4), w, b, 4)
xq = Q.functions.DorefaQuant(xu.to(dev), 4)
tag = xq._qt_codes
print("xq eq", torch.equal(xq.cpu(), O.dorefa_quantize(xu,4)))
ca = O.dorefa_act_codes(xu,4)
print("codes eq", np.array_equal(tag.codes.cpu().numpy()[:, :512].astype(np.int64), ca), "scale", tag.scale, "rowsum eq", np.array_equal(tag.row_sum.cpu().numpy(), ca.sum(1)))
p = ops.pack_weight(w.to(dev), "dorefa", 4, want_wq=True)
print("stats", p.stats.cpu()[:4], "maxtanh ref", torch.tanh(w).abs().max().item())
cw = O.dorefa_weight_codes(w,4)
ws, ldw = ops.expand_weight(p, L.CODES_I8)
got = ws.cpu().numpy()[:, :512].astype(np.int64)
print("wcodes eq", np.array_equal(got, 2*cw-15), "ndiff", (got != 2*cw-15).sum(), got[0,:8], (2*cw-15)[0,:8])
print("wq close", (p.wq.cpu()-O.dorefa_weight(w,4)).abs().max().item(), "col_scale", p.col_scale[:3].cpu())
for be in ("simt","tcgen05"):
    Q.set_backend(i8=be)
    yd = ld.to(dev)(xq).cpu()
    print(be, "rel", float((yd - refd).abs().max() / refd.abs().max()))
Q.set_backend(i8="auto")

# ---- kernel microbench with CUDA events
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); s.record()
    for _ in range(iters): fn()
    e.record(); t1=time.perf_counter(); torch.cuda.synchronize()
    return s.elapsed_time(e)/iters, (t1-t0)/iters*1e3
M,K,N=8192,4096,4096
x=torch.randn(M,K,device=dev); wf=torch.randn(N,K,device=dev)*0.02
_, ta = ops.quant_act(x, L.Q_SIGN, want_y=False, codes_kind=L.CODES_I8, want_bits=True, kind="sign")
pw = ops.pack_weight(wf,"sign")
wsx, ldwx = ops.expand_weight(pw, L.CODES_I8)
out=torch.empty(M,N,device=dev)
epi=ops.make_epi(out, ldo=N)
print("gemm_i8 tc  gpu_ms,host_ms:", timeit(lambda: ops.gemm_i8(ta.codes, True, ta.ld, wsx, True, ldwx, M,N,K, epi, L.BACKEND_TCGEN05)))
epi2=ops.make_epi(None, ldo=N, acc_out=torch.empty(M,N,dtype=torch.int32,device=dev))
print("gemm_i8 tc acc_out only:", timeit(lambda: ops.gemm_i8(ta.codes, True, ta.ld, wsx, True, ldwx, M,N,K, epi2, L.BACKEND_TCGEN05)))
print("expand w i8:", timeit(lambda: ops.expand_weight(pw, L.CODES_I8)))
print("quant sign y+codes+bits:", timeit(lambda: ops.quant_act(x, L.Q_SIGN, want_y=True, codes_kind=L.CODES_I8, want_bits=True, kind="sign")))
print("quant sign codes only:", timeit(lambda: ops.quant_act(x, L.Q_SIGN, want_y=False, codes_kind=L.CODES_I8, kind="sign")))
print("quant xnor y+bf16:", timeit(lambda: ops.quant_act(x, L.Q_XNOR_ROW, want_y=True, codes_kind=L.CODES_BF16, want_row_scale=True, kind="xnor")))
print("torch copy 134MB:", timeit(lambda: out.copy_(x)))
lay = Q.layers.LinearBin(K,N).to(dev).eval(); act=Q.functions.BinaryConnect()
with torch.no_grad():
    xq2=act(x)
    print("layer(xq) eval:", timeit(lambda: lay(xq2)))
    print("act+layer eval:", timeit(lambda: lay(act(x))))
print("popcount b1b1:", timeit(lambda: ops.gemm_b1b1(ta.bits, ta.ld_bits, pw.packed.view(torch.int32)[0], pw.ld_packed//4, M,N,K, epi), iters=5))
