import sys, torch
sys.path.insert(0, '/root/repo')
import pytorch_quantize_impls_b200 as Q
dev='cuda'
M,K,N=8192,4096,4096
torch.manual_seed(0)
x=torch.randn(M,K,device=dev)
with torch.no_grad():
    lay=Q.layers.LinearBin(K,N).to(dev).eval(); act=Q.functions.BinaryConnect()
    xq=act(x)
    for _ in range(3): y=lay(xq)
    lx=Q.layers.LinearXNOR(K,N).to(dev).eval(); xx=Q.functions.QuantXnor(x,1)
    for _ in range(3): y=lx(xx)
torch.cuda.synchronize()
