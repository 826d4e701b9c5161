import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lct-gan_b200")
for p in (PKG, ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "bf16: run with the tcgen05 bf16 tensor-core path enabled (default for tests: fp32)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def dev():
    import torch
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _precision_mode(request):
    """GPU parity tests run the fp32 kernels unless marked `bf16` (then the tcgen05 dense path is on)."""
    if "gpu" not in request.keywords:
        yield
        return
    from lctgan import config
    config.set_precision("bf16" if request.node.get_closest_marker("bf16") else "fp32")
    yield
    config.set_precision("bf16")
