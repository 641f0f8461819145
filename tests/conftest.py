"""pytest config: registers the `gpu` marker and puts the repo root + the product
package directory (torch-mnf_b200/, which holds the drop-in `torch_mnf` package) on sys.path."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "torch-mnf_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _inference_mode_by_default(request):
    """Parity tests exercise the inference kernels; modules whose name contains 'autograd' keep grad enabled
    (with grad enabled and trainable parameters the flows take the differentiable path, like the reference)."""
    import torch

    keep = "autograd" in request.module.__name__ or "train" in request.module.__name__
    with torch.set_grad_enabled(keep):
        yield
