import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_library_is_built():
    """The C-ABI library is git-ignored (built in-tree by __graft_entry__.build()): make sure a fresh checkout has a
    current one before any test loads it.  A no-op (a digest comparison) when it is already up to date."""
    import importlib.util
    path = os.path.join(ROOT, "restrictive-hierarchical-semantic-segmentation_b200", "build.py")
    spec = importlib.util.spec_from_file_location("_rhseg_build", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not mod.is_current():
        mod.build()
    yield
