"""Generate tests/golden/quanttorch_ref_v1.npz from the LIVE reference.

TEST INFRASTRUCTURE.  Run in the build container (where /root/reference
exists):   python oracle/gen_golden.py

Every array under key ``<case>/out*`` is an output of the unmodified reference
code (QuantTorch.functions / QuantTorch.layers, imported through
oracle/ref_loader.py); ``<case>/x``, ``/w``, ``/b`` are the seeded inputs it was
run on.  The committed .npz is what the GPU box sees (it has no
/root/reference).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "quanttorch_ref_v1.npz")


def special_values():
    return torch.tensor([0.0, -0.0, float("nan"), 1e-45, -1e-45, 0.5, -0.5, 0.25, -0.25, 0.75, -0.75,
                         1.0, -1.0, 1.001, 1.002, -1.002, 3.0, -3.0, 1 / 6, 0.1667, 0.8333, 1.2, -0.3],
                        dtype=torch.float32)


def main():
    ref = load_reference()
    if ref is None:
        raise SystemExit("reference tree not found")
    Fn, L = ref
    torch.manual_seed(20240917)
    torch.set_num_threads(1)
    g = {}

    def put(name, **arrs):
        for k, v in arrs.items():
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            g[f"{name}/{k}"] = np.asarray(v)

    # ---------------- elementwise ops ----------------
    sv = special_values()
    rnd = torch.cat([sv, torch.empty(200).uniform_(-1.5, 1.5)])
    put("safe_sign", x=rnd, out=Fn.safeSign(rnd))
    put("binary_det", x=rnd, out=Fn.BinaryConnectDeterministic.apply(rnd))
    put("ternary_det", x=rnd, out=Fn.TernaryConnectDeterministic.apply(rnd))
    # reference KATs, tests/implementations/Terner/function_test.py:10-27
    k1 = torch.tensor([0.75, 0.5, 0.25, 0, -1, -0.2])
    k2 = torch.tensor([1, 0, .51, .1, 0, -1, -.2, .7])
    put("ternary_kat1", x=k1, out=Fn.TernaryConnectDeterministic.apply(k1))
    put("ternary_kat2", x=k2, out=Fn.TernaryConnectDeterministic.apply(k2))
    nonan = rnd[~torch.isnan(rnd)]
    half = torch.tensor([(i + 0.5) / 15 for i in range(-3, 18)] + [(i + 0.5) / 3 for i in range(-2, 5)])
    dq_in = torch.cat([nonan, half, torch.empty(100).uniform_(0, 1)])
    for k in (1, 2, 3, 4, 5, 8, 32):
        put(f"dorefa_quant_k{k}", x=dq_in, out=Fn.DorefaQuant(dq_in, k))
    wmat = torch.empty(24, 70).uniform_(-0.9, 0.9)
    wmat[0, 0] = 0.0
    for k in (1, 2, 3, 4, 5, 8, 32):
        put(f"dorefa_weight_k{k}", w=wmat, out=Fn.nnQuantWeight(k)(wmat))
    put("dorefa_weight_zero", w=torch.zeros(3, 5), out=Fn.nnQuantWeight(3)(torch.zeros(3, 5)))
    xa = torch.randn(33, 70)
    xa[1, 2] = 0.0
    for d in (-1, 0, 1):
        put(f"xnor_act_dim{d}", x=xa, out=Fn.QuantXnor(xa, d))
    lg_in = torch.cat([torch.tensor([0.3, 1.234, 5, -1, 0, 200.0, -1e-3]), torch.empty(100).uniform_(-130, 130)])
    for fsr, bw in ((7, 3), (2, 2), (5, 4)):
        put(f"log_quant_{fsr}_{bw}", x=lg_in, out=Fn.Quant(lg_in, "log", fsr, bw))
        put(f"lin_quant_{fsr}_{bw}", x=lg_in, out=Fn.Quant(lg_in, "lin", fsr, bw))
    put("lin_quant_unsigned_2_8", x=torch.tensor([0.3, 1.234, 5, -1]),
        out=Fn.Quant(torch.tensor([0.3, 1.234, 5, -1]), "lin", 2, 8, with_sign=False))

    # ---------------- dense layers (train-mode forward) ----------------
    M, K, N = 33, 70, 24
    x = torch.randn(M, K)
    x[0, :4] = torch.tensor([0.0, -0.0, 0.5, -0.5])
    xu = torch.empty(M, K).uniform_(0, 1)
    b = torch.empty(N).uniform_(-1, 1)

    def set_wb(layer, w):
        layer.weight.data.copy_(w)
        layer.bias.data.copy_(b)

    wn = torch.randn(N, K) * (1.0 / K) ** 0.5 * 3      # spreads past +-0.5 so ternary has all 3 levels
    wn[0, :3] = torch.tensor([0.0, 0.5, -0.5])
    lay = L.LinearBin(K, N); set_wb(lay, wn)
    xb = Fn.BinaryConnectDeterministic.apply(x)
    put("lin_bin", x=x, xq=xb, w=wn, b=b, out=lay(xb), out_real=lay(x))
    lay.train(False)
    put("lin_bin_eval", w_eval=lay.weight.data, out=lay(xb))
    lay = L.LinearTer(K, N); set_wb(lay, wn)
    put("lin_ter", x=x, xq=xb, w=wn, b=b, out=lay(xb), out_real=lay(x))
    for k in (1, 2, 3, 4, 8):
        lay = L.LinearDorefa(K, N, bit_width=k); set_wb(lay, wmat)
        for ka in (k, 8 if k != 8 else 4):
            xq = Fn.DorefaQuant(xu, ka)
            put(f"lin_dorefa_w{k}a{ka}", x=xu, xq=xq, w=wmat, b=b, out=lay(xq))
        put(f"lin_dorefa_w{k}_real", x=x, w=wmat, b=b, out=lay(x))
    lay = L.LinearXNOR(K, N); set_wb(lay, wmat)
    xq = Fn.QuantXnor(xa, 1)
    put("lin_xnor", x=xa, xq=xq, w=wmat, b=b, out=lay(xq), out_real=lay(xa))
    for dt, fsr, bw in (("lin", 7, 3), ("log", 7, 3), ("log", 2, 2)):
        lay = L.LinearQuant(K, N, dtype=dt, fsr=fsr, bit_width=bw)
        wl = torch.empty(N, K).uniform_(-2 ** fsr, 2 ** fsr)
        set_wb(lay, wl)
        put(f"lin_loglin_{dt}_{fsr}_{bw}", x=x, w=wl, b=b, out=lay(x))

    # ---------------- conv layers ----------------
    xi = torch.randn(2, 5, 9, 9)
    xiu = torch.empty(2, 5, 9, 9).uniform_(0, 1)
    wc = torch.randn(7, 5, 3, 3) * 0.5
    wc[0, 0, 0, :] = torch.tensor([0.0, 0.5, -0.5])
    bc = torch.empty(7).uniform_(-1, 1)
    for tag, kw in (("s1p1", dict(stride=1, padding=1)), ("s2p0", dict(stride=2, padding=0)),
                    ("s1p2d2", dict(stride=1, padding=2, dilation=2))):
        lay = L.BinConv2d(5, 7, 3, **kw); lay.weight.data.copy_(wc); lay.bias.data.copy_(bc)
        xq = Fn.BinaryConnectDeterministic.apply(xi)
        put(f"conv_bin_{tag}", x=xi, xq=xq, w=wc, b=bc, out=lay(xq), out_real=lay(xi))
        lay = L.TerConv2d(5, 7, 3, **kw); lay.weight.data.copy_(wc); lay.bias.data.copy_(bc)
        put(f"conv_ter_{tag}", x=xi, xq=xq, w=wc, b=bc, out=lay(xq))
        for k in (2, 4, 8):
            lay = L.DorefaConv2d(5, 7, 3, bit_width=k, **kw); lay.weight.data.copy_(wc); lay.bias.data.copy_(bc)
            xq4 = Fn.DorefaQuant(xiu, k)
            put(f"conv_dorefa_w{k}a{k}_{tag}", x=xiu, xq=xq4, w=wc, b=bc, out=lay(xq4))
        lay = L.XNORConv2d(5, 7, 3, **kw); lay.weight.data.copy_(wc); lay.bias.data.copy_(bc)
        put(f"conv_xnor_{tag}", x=xi, w=wc, b=bc, out=lay(xi))
    # grouped conv
    wg = torch.randn(8, 3, 3, 3)
    xg = torch.randn(2, 6, 7, 7)
    lay = L.BinConv2d(6, 8, 3, padding=1, groups=2, bias=False); lay.weight.data.copy_(wg)
    xq = Fn.BinaryConnectDeterministic.apply(xg)
    put("conv_bin_groups2", x=xg, xq=xq, w=wg, out=lay(xq))

    # reference layer KATs: tests/implementations/BinaryNet/layer_test.py:16-21
    lay = L.LinearBin(3, 1, bias=False)
    lay.weight.data.copy_(torch.tensor([[0.5, 0.0, -0.5]]))
    xk = torch.tensor([[2.0, 1.0, -3.0]])
    put("lin_bin_kat", x=xk, w=lay.weight.data, out=lay(xk))

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, len(g), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
