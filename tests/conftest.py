import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "experimental: variants compiled but not yet run on a GPU; skipped unless "
                            "EK_TEST_EXPERIMENTAL=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("EK_TEST_EXPERIMENTAL") == "1":
        return
    skip = pytest.mark.skip(reason="experimental variant: set EK_TEST_EXPERIMENTAL=1")
    for item in items:
        if "experimental" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CUDA library, the host library and the oracle are built (cheap if up to date)."""
    import __graft_entry__ as g

    g.build(quiet=True)
    yield
