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


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not errored) on a box without a CUDA device, whatever -m says."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
