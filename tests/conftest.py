import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of failing them (the product path has no
    CPU fallback, so they cannot run there)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked tests run on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """Reference-produced golden vectors (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "ljspeech_world_golden.npz"))


def golden_utterance(golden, id_):
    """x (pre-emphasised 0.97, fp64), cmp [T x 67], f0 [T] (exp of cmp col 60 where col 63 is voiced), fs."""
    x = golden[id_ + "/wav"].astype(np.float64) / 32768.0
    x = np.append(x[0], x[1:] - 0.97 * x[:-1])
    c = golden[id_ + "/cmp"]
    f0 = np.where(c[:, 63] > 0, np.exp(c[:, 60].astype(np.float64)), 0.0)
    return x, c, f0, 16000
