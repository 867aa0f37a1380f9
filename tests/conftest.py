import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "quanttorch_ref_v1.npz")
    z = np.load(path)
    return {k: z[k] for k in z.files}


def case(golden, name):
    """Return dict of torch tensors for golden case `name`."""
    import torch
    pre = name + "/"
    return {k[len(pre):]: torch.from_numpy(v.copy()) for k, v in golden.items() if k.startswith(pre)}
