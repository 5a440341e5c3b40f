import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "audio-rag_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


REFERENCE_SRC = "/root/reference/src"
HAVE_REFERENCE = os.path.isdir(os.path.join(REFERENCE_SRC, "audio_rag"))
if HAVE_REFERENCE:
    # Make the reference's own plugin code importable (read-only use; never at GPU-test time, where it is absent):
    #   * a stub `sentence_transformers` (reranking/bge.py:3-4 imports it at module top)
    #   * tests/fake_qdrant: a test double of the third-party qdrant_client backed by the oracle
    import types
    if "sentence_transformers" not in sys.modules:
        st = types.ModuleType("sentence_transformers")
        st.CrossEncoder = st.SentenceTransformer = object
        sys.modules["sentence_transformers"] = st
    sys.path.insert(0, os.path.join(ROOT, "tests", "fake_qdrant"))
    sys.path.append(REFERENCE_SRC)


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
