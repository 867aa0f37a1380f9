"""Generate tests/golden/quanttorch_ref_loglin_v1.npz from the LIVE reference: LogLin layers fed by Lin / Log quantized
activations (the chains the k-bit LogLin format of SURVEY.md 8f-3 accelerates).

TEST INFRASTRUCTURE.  Run in the build container (where /root/reference exists):   python oracle/gen_golden_loglin.py
Keys: <case>/x (seeded input), /xq (reference activation quantizer output), /w, /b, /out (reference layer output)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "quanttorch_ref_loglin_v1.npz")


def main():
    ref = load_reference()
    if ref is None:
        raise SystemExit("reference tree not found")
    Fn, L = ref
    torch.manual_seed(20241017)
    torch.set_num_threads(1)
    g = {}

    def put(name, **arrs):
        for k, v in arrs.items():
            g[f"{name}/{k}"] = np.asarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v)

    with torch.no_grad():
        # dense: activation quantizer (dtype_a, fsr_a, bw_a) -> LinearQuant(dtype_w, fsr_w, bw_w)
        for (da, fa, ba), (dw, fw, bw) in [(("lin", 2, 4), ("lin", 0, 4)), (("lin", 1, 6), ("lin", -1, 3)),
                                           (("log", 1, 3), ("log", 1, 3)), (("lin", 0, 2), ("log", 0, 2)),
                                           (("log", 2, 2), ("lin", 2, 5))]:
            lay = L.LinearQuant(72, 40, dtype=dw, fsr=fw, bit_width=bw)
            lay.bias.data.uniform_(-1, 1)
            x = torch.randn(33, 72) * (2.0 ** fa) * 0.6
            x[0, :4] = torch.tensor([0.0, 2.0 ** fa, -(2.0 ** fa) * 3, 2.0 ** (fa - ba) * 0.5])
            xq = Fn.Quant(x, da, fa, ba, True)
            put(f"dense_{da}{fa}_{ba}__{dw}{fw}_{bw}", x=x, xq=xq, w=lay.weight.data, b=lay.bias.data, out=lay(xq))
        # conv (training mode: the reference quantizes the conv weights in train mode only, log_lin_layers.py:87-93)
        for (da, fa, ba), (dw, fw, bw), kw in [(("lin", 1, 4), ("lin", 0, 3), dict(padding=1)),
                                               (("lin", 0, 3), ("log", 1, 3), dict(stride=2, padding=1))]:
            conv = L.QuantConv2d(32, 24, 3, fsr=fw, bit_width=bw, dtype=dw, **kw)
            conv.bias.data.uniform_(-1, 1)
            conv.train()
            x = torch.randn(2, 32, 7, 7) * (2.0 ** fa) * 0.6
            xq = Fn.Quant(x, da, fa, ba, True)
            put(f"conv_{da}{fa}_{ba}__{dw}{fw}_{bw}", x=x, xq=xq, w=conv.weight.data, b=conv.bias.data, out=conv(xq),
                stride=np.int64(kw.get("stride", 1)), padding=np.int64(kw.get("padding", 0)))
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, len(g), "arrays")


if __name__ == "__main__":
    main()
