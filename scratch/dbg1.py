import sys, os, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import quanttorch_oracle as O
import pytorch_quantize_impls_b200 as Q
from pytorch_quantize_impls_b200 import _lib as L, _ops as ops
torch.manual_seed(0)
M,K,N=256,512,384
x=torch.randn(M,K); w=torch.randn(N,K)*0.1; b=torch.rand(N)
_, tag = ops.quant_act(x.cuda(), L.Q_SIGN, want_y=False, codes_kind=L.CODES_I8, want_bits=True, kind="sign")
codes=tag.codes.cpu().numpy()[:, :K].astype(np.int64)
print("codes ok", np.array_equal(codes, O.sign_codes(x)), tag.ld)
p=ops.pack_weight(w.cuda(),"sign")
ws,ldw=ops.expand_weight(p,L.CODES_I8)
print("w expand ok", np.array_equal(ws.cpu().numpy()[:, :K].astype(np.int64), O.sign_codes(w)), ldw)
for be in (L.BACKEND_SIMT, L.BACKEND_TCGEN05):
    acc=torch.zeros(M,N,dtype=torch.int32).cuda(); out=torch.zeros(M,N).cuda()
    ops.gemm_i8(tag.codes, True, tag.ld, ws, True, ldw, M,N,K, ops.make_epi(out, ldo=N, acc_out=acc, bias=b.cuda()), be)
    ref=O.int_acc(O.sign_codes(x),O.sign_codes(w))
    a=acc.cpu().numpy()
    bad=np.argwhere(a!=ref)
    print("backend",be,"acc ok",len(bad)==0, "nbad",len(bad), bad[:5], a[tuple(bad[0])] if len(bad) else None, ref[tuple(bad[0])] if len(bad) else None)
    yref=torch.from_numpy(ref).float()+b
    print(" out ok", torch.equal(out.cpu(), yref), (out.cpu()-yref).abs().max().item())
    bad=(out.cpu()!=yref).nonzero()
    print(bad[:10].tolist())
