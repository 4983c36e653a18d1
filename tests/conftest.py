import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def meshes():
    d = np.load(os.path.join(GOLDEN, "meshes.npz"))
    names = sorted({k.rsplit("_", 1)[0] for k in d.files})
    return {n: (d[f"{n}_verts"], d[f"{n}_tets"].astype(np.int64)) for n in names}


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))
