"""Generate tests/golden/quanttorch_ref_functional_v1.npz from the LIVE reference: the *functional* dense / conv ops
(SURVEY.md 8a rows a5, a14, a20), forward outputs and the gradients of their hand-written backward.

TEST INFRASTRUCTURE.  Run in the build container (where /root/reference exists):
    python oracle/gen_golden_functional.py

    a5   BinaryDense                                     QuantTorch/functions/binary_connect.py:86-112
    a14  TernaryDense / TernaryConv2d (deterministic)    QuantTorch/functions/terner_connect.py:78-153
         (torch.sign semantics: +-0.5 -> +-0.5, 0 -> 0 -- NOT the layer path's safeSign thresholds)
    a20  QuantDense / QuantConv2d                        QuantTorch/functions/dorefa_connect.py:116-199

Keys: <case>/x, /w, /b, /go (seeded inputs and the output gradient), /out, /gx, /gw, /gb (reference results).
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "quanttorch_ref_functional_v1.npz")


def main():
    ref = load_reference()
    if ref is None:
        raise SystemExit("reference tree not found")
    Fn, _ = ref
    # functions/__init__.py:25 re-binds the name QuantConv2d to the Elastic op; the DoReFa one is reached through its module
    from QuantTorch.functions import dorefa_connect as DF
    torch.manual_seed(20241017)
    torch.set_num_threads(1)
    g = {}

    def run(name, op, x, w, b):
        x = x.clone().requires_grad_(True)
        w = w.clone().requires_grad_(True)
        b = None if b is None else b.clone().requires_grad_(True)
        y = op.apply(x, w, b) if b is not None else op.apply(x, w)
        go = torch.randn(y.shape)
        y.backward(go)
        arrs = dict(x=x, w=w, go=go, out=y, gx=x.grad, gw=w.grad)
        if b is not None:
            arrs.update(b=b, gb=b.grad)
        for k, v in arrs.items():
            g[f"{name}/{k}"] = v.detach().cpu().numpy()

    M, K, N = 37, 70, 24
    x = torch.randn(M, K)
    x[0, :4] = torch.tensor([0.0, -0.0, 0.5, -0.5])
    xu = torch.empty(M, K).uniform_(0, 1)
    b = torch.empty(N).uniform_(-1, 1)
    w = torch.randn(N, K) * 0.6
    w[0, :6] = torch.tensor([0.0, -0.0, 0.5, -0.5, 0.25, -0.75])        # ties and zeros of the torch.sign ternary form
    wsmall = torch.empty(N, K).uniform_(-0.9, 0.9)
    wsmall[1, 0] = 0.0

    run("binary_dense", Fn.BinaryDense, x, w, b)
    run("binary_dense_nobias", Fn.BinaryDense, x, w, None)
    run("binary_dense_quant_in", Fn.BinaryDense, Fn.BinaryConnectDeterministic.apply(x).detach(), w, b)
    run("ternary_dense", Fn.TernaryDense(False), x, w, b)
    run("ternary_dense_nobias", Fn.TernaryDense(False), x, w, None)
    for k in (1, 2, 3, 4, 32):
        run(f"quant_dense_k{k}", DF.QuantDense(k), xu, wsmall, b)
    run("quant_dense_k3_real", DF.QuantDense(3), x, wsmall, None)

    xi = torch.randn(2, 5, 9, 9)
    xiu = torch.empty(2, 5, 9, 9).uniform_(0, 1)
    wc = torch.randn(7, 5, 3, 3) * 0.5
    wc[0, 0, 0, :] = torch.tensor([0.0, 0.5, -0.5])
    bc = torch.empty(7).uniform_(-1, 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for tag, kw in (("s1p1", dict(stride=1, padding=1)), ("s2p0", dict(stride=2, padding=0))):
            run(f"ternary_conv_{tag}", Fn.TernaryConv2d(False, **kw), xi, wc, bc)
            for k in (1, 2, 3):
                run(f"quant_conv_k{k}_{tag}", DF.QuantConv2d(bit_width=k, **kw), xiu, wc, bc)
        run("ternary_conv_nobias", Fn.TernaryConv2d(False, stride=1, padding=1), xi, wc, None)

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, len(g), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
