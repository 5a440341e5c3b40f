import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "audio-rag_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Make sure libb200rag.so and the C oracle exist (compiles here on CPU; prebuilt files travel to the GPU box)."""
    from b200rag import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        sys.path.insert(0, os.path.join(ROOT, "audio-rag_b200"))
        import build as _b
        _b.build()
    return _ffi.load()


@pytest.fixture(scope="session")
def gpu(built_lib):
    from b200rag import _ffi
    if _ffi.device_count() < 1:
        pytest.fail("no sm_100 device visible: -m gpu tests must run on a B200 (no CPU fallback exists)")
    return 0
