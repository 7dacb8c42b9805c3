import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return {k: (torch.from_numpy(z[k]) if (z[k].ndim > 0 or k.startswith('sd.')) else z[k].item()) for k in z.files}
