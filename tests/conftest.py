import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "blockcopy-video-processing-pytorch_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The C-ABI library must exist before anything imports blockcopy._C (nvcc cross-compiles
    without a GPU); the CPU oracle is built alongside."""
    sys.path.insert(0, PKG)
    import importlib.util

    spec = importlib.util.spec_from_file_location("bc_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(mod.OUT) or mod._stale():
        mod.build_library()
    from oracle import cpu_oracle

    cpu_oracle.build()
    yield


@pytest.fixture
def golden_dir():
    return GOLDEN
