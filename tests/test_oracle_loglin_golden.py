"""CPU: the oracle's LogLin restatement (lin_quant / log_quant / linear_loglin / conv_loglin) reproduces, bit for bit, what the
LIVE reference computed for LogLin layers fed by Lin / Log quantized activations (tests/golden/quanttorch_ref_loglin_v1.npz,
written by oracle/gen_golden_loglin.py).  These chains are what the k-bit LogLin format accelerates; the GPU
twin of this test is tests/test_gpu_loglin.py::test_loglin_chains_match_live_reference_goldens."""
import os
import re

import numpy as np
import pytest
import torch

import quanttorch_oracle as O

PATH = os.path.join(os.path.dirname(__file__), "golden", "quanttorch_ref_loglin_v1.npz")
NAME = re.compile(r"(dense|conv)_(lin|log)(-?\d+)_(\d+)__(lin|log)(-?\d+)_(\d+)")


def load_cases():
    z = np.load(PATH)
    cases = {}
    for k in z.files:
        name, field = k.split("/")
        cases.setdefault(name, {})[field] = torch.from_numpy(z[k].copy()) if z[k].ndim else int(z[k])
    return cases


CASES = load_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_live_reference(name):
    c = CASES[name]
    kind, da, fa, ba, dw, fw, bw = NAME.fullmatch(name).groups()
    fa, ba, fw, bw = int(fa), int(ba), int(fw), int(bw)
    xq = (O.lin_quant if da == "lin" else O.log_quant)(c["x"], fa, ba, True)
    assert torch.equal(xq, c["xq"])
    if kind == "dense":
        out = O.linear_loglin(xq, c["w"], c["b"], dw, fw, bw)
    else:
        out = O.conv_loglin(xq, c["w"], c["b"], dw, fw, bw, stride=c["stride"], padding=c["padding"])
    assert torch.equal(out, c["out"])
