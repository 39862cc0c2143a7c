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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def sub_state(g, prefix):
    """Tensors of a golden dict whose key starts with prefix, prefix stripped."""
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


@pytest.fixture(scope="session")
def office_build():
    return load_golden("office_a2d_build.npz")


@pytest.fixture(scope="session")
def office_mp():
    return load_golden("office_a2d_mp.npz")


@pytest.fixture(scope="session")
def fb_build():
    return load_golden("fb_h2c_cosine_build.npz")


@pytest.fixture(scope="session")
def diag_golden():
    return load_golden("diagnostics.npz")


@pytest.fixture(scope="session")
def office_assemble():
    return load_golden("office_a2d_assemble.npz")
