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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden(dict):
    """npz fixture -> dict of torch tensors (0-d arrays become python scalars)."""

    def __getattr__(self, k):
        return self[k]


def load_golden(name):
    raw = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = Golden()
    for k in raw.files:
        a = raw[k]
        out[k] = a.item() if a.ndim == 0 else torch.from_numpy(a)
    return out


@pytest.fixture
def golden():
    return load_golden


def assert_close(a, b, rtol=1e-4, atol=1e-5, what=""):
    """|a-b| <= rtol*|b| + atol elementwise (SURVEY.md section 8d parity metric)."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    err = (a - b).abs()
    tol = rtol * b.abs() + atol
    bad = err > tol
    if bad.any():
        i = torch.nonzero(bad)[0].tolist()
        worst = (err - tol).argmax()
        raise AssertionError("%s: %d/%d elements out of tolerance; first at %s got %r want %r; worst err %.3e"
                             % (what, int(bad.sum()), bad.numel(), i, a[tuple(i)].item(), b[tuple(i)].item(),
                                err.flatten()[worst].item()))
