import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import fast_dnn_b200  # noqa: E402,F401  (alias of the hyphenated package directory)
from fast_dnn_b200 import synth  # noqa: E402

import oracle_py  # noqa: E402  TEST INFRASTRUCTURE: the CPU checker

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_ROOT = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    # building the checker is not using it; the reference build only happens where /root/reference exists
    oracle_py.build(port=True, ref=os.path.isdir(REFERENCE_ROOT))


@pytest.fixture(scope="session")
def net_file():
    def make(shape, seed=1234, stress=False):
        return synth.network_file(shape, seed=seed, stress=stress)
    return make


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir(REFERENCE_ROOT) and oracle_py.have_ref()


def softmax_close(got: np.ndarray, want: np.ndarray):
    """The stated float tolerance on softmax scores: the logits are
    bit-exact; the scores differ only through expf (CUDA ≤ 2 ulp vs glibc) and the order of the
    fp32 sum of exponentials (tree vs the reference's sequential loop)."""
    tol = 1e-9 + 2e-5 * np.abs(want)
    bad = np.abs(got.astype(np.float64) - want.astype(np.float64)) > tol
    assert not bad.any(), f"{bad.sum()} of {bad.size} scores outside |d| <= 1e-9 + 2e-5*|ref|; worst rel " \
                          f"{np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-30)):.3e}"
