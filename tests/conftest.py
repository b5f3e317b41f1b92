import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree; build it if a fresh checkout has not yet (nvcc works without a GPU)."""
    import magi_b200  # noqa: F401
    from magi_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    elif not _lib.binary_is_current():
        # a binary built from other sources than the tree's (source stamp in magi_build_info): rebuild before any test
        # loads it; the loader caches the handle, so this must happen in a fresh process state
        _lib._lib = None
        _lib.build()
        assert _lib.binary_is_current(), "libmagi_attn.so is stale and could not be rebuilt"
    return _lib.LIB_PATH
