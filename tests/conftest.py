import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def k1_golden():
    z = np.load(GOLDEN / "k1_backproject.npz")
    meta = json.loads(bytes(z["meta_json"]).decode())
    return z, meta


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test ran without a CUDA device: there is no CPU fallback")
    from textureless_3d_reconstruction_b200.runtime import get_context
    return get_context(0)


@pytest.fixture(scope="session")
def oracle():
    from oracle import capi
    capi.lib()
    return capi
