import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_weights(name):
    """tests/golden/weights_<name>.npz -> (state_dict of torch tensors, config dict)."""
    import torch

    z = np.load(os.path.join(GOLDEN, f"weights_{name}.npz"))
    cfg = json.loads(bytes(z["__config__"]).decode())
    sd = {k: torch.from_numpy(z[k].copy()) for k in z.files if k != "__config__"}
    return sd, cfg


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        return np.load(os.path.join(GOLDEN, name))

    return _load


CHECKPOINTS = ["dari_tult", "dari_tult2", "good"]
