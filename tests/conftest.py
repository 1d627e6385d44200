import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes minutes on the CPU; set CROPSR_SLOW=1 to run")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("CROPSR_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="slow; set CROPSR_SLOW=1")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def manifest():
    import json
    with open(os.path.join(GOLDEN, "cases", "manifest.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library; building it is the job of __graft_entry__.build()."""
    import __graft_entry__
    if not os.path.exists(os.path.join(ROOT, "cropsr_b200", "libcropsr_b200.so")):
        __graft_entry__.build()
    from cropsr_b200 import _native
    return _native
