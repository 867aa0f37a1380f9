"""Golden gradient vectors from the LIVE reference (build container only): tests/golden/quanttorch_ref_grads_v1.npz.
For each family: y = layer(act(x)); loss = sum(y * g); grads w.r.t. x, W, b through the reference's own STE backward
(binary_connect.py:30-38, terner_connect.py:29-34, dorefa_connect.py:41-44,66-79, xnor_connect.py:30-37,118-130).
TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "quanttorch_ref_grads_v1.npz")


def main():
    Fn, L = load_reference()
    torch.manual_seed(77)
    g = {}

    def run(name, layer, act, x, w, b):
        layer.weight.data.copy_(w)
        if b is not None:
            layer.bias.data.copy_(b)
        xr = x.clone().requires_grad_(True)
        y = layer(act(xr) if act is not None else xr)
        go = torch.randn_like(y)
        (y * go).sum().backward()
        g[name + "/x"], g[name + "/w"], g[name + "/go"] = x.numpy(), w.numpy(), go.numpy()
        g[name + "/y"] = y.detach().numpy()
        g[name + "/gx"], g[name + "/gw"] = xr.grad.numpy(), layer.weight.grad.numpy()
        if b is not None:
            g[name + "/b"], g[name + "/gb"] = b.numpy(), layer.bias.grad.numpy()

    M, K, N = 19, 40, 12
    x = torch.randn(M, K) * 0.9
    x[0, :3] = torch.tensor([1.0005, 1.002, -1.5])             # STE clip boundary |x| <= 1.001
    w = torch.randn(N, K) * 0.7
    w[0, :3] = torch.tensor([1.0005, 1.002, -1.2])
    b = torch.rand(N) - 0.5
    xu = torch.rand(M, K)
    run("lin_bin", L.LinearBin(K, N), Fn.BinaryConnect(), x, w, b)
    run("lin_ter", L.LinearTer(K, N), Fn.BinaryConnect(), x, w, b)
    for k in (1, 2, 4):
        run(f"lin_dorefa{k}", L.LinearDorefa(K, N, bit_width=k), Fn.nnDorefaQuant(k), xu, w, b)
    run("lin_xnor", L.LinearXNOR(K, N), Fn.nnQuantXnor(1), x, w, b)
    xi = torch.randn(2, 4, 7, 7) * 0.9
    wc = torch.randn(6, 4, 3, 3) * 0.7
    bc = torch.rand(6) - 0.5
    run("conv_bin", L.BinConv2d(4, 6, 3, stride=2, padding=1), Fn.BinaryConnect(), xi, wc, bc)
    run("conv_ter", L.TerConv2d(4, 6, 3, padding=1), Fn.BinaryConnect(), xi, wc, bc)
    run("conv_dorefa3", L.DorefaConv2d(4, 6, 3, padding=1, bit_width=3), Fn.nnDorefaQuant(3), xi.abs().clamp(0, 1), wc, bc)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, len(g))


if __name__ == "__main__":
    main()
