import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    """On a machine without CUDA the `gpu` tests are skipped (so any `pytest tests/...` selection is a
    clean CPU run) -- unless they were asked for by name with `-m gpu`: then a missing device must fail
    loudly, which is what the round-end run on the B200 box relies on."""
    import torch
    expr = (config.getoption("-m") or "").strip()
    if torch.cuda.is_available() or expr == "gpu":
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run with `-m gpu` on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    import torch

    def load(name):
        with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
            return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}

    return load
