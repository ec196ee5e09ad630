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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True, scope="session")
def _three_term_precision_unless_a_test_says_otherwise():
    """The parity tests against the fp64 oracle / the reference's golden vectors pin the fp32-equivalent (three-term) arithmetic;
    the fp16-operand modes have their own tests (test_gpu_fwd_fp16.py, test_gpu_mlp_bwd.py) which select them explicitly.
    Subprocess tests (the reference scripts through the drop-in) run with the package defaults."""
    if not torch.cuda.is_available():
        yield
        return
    from consistentnerf_b200 import ops
    prev = (ops.set_forward_precision("split"), ops.set_grad_precision("split"))
    yield
    ops.set_forward_precision(prev[0])
    ops.set_grad_precision(prev[1])


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def t(a, dtype=None, device=None):
    x = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        x = x.to(dtype)
    if device is not None:
        x = x.to(device)
    return x
