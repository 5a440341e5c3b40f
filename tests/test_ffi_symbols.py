"""The C-ABI library loads without a GPU, exports every symbol include/b200rag.h declares, and refuses compute
without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound(built_lib):
    from b200rag import _ffi
    hdr = open(os.path.join(ROOT, "include", "b200rag.h")).read()
    declared = set(re.findall(r"\b(b200rag_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b200rag_shard", "b200rag_config", "b200rag_query", "b200rag_cand", "b200rag_stats"}
    assert len(declared) >= 32
    bound = {n for n, _, _ in _ffi.SYMBOLS}
    assert declared == bound, f"header/binding mismatch: {declared ^ bound}"
    raw = ctypes.CDLL(_ffi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name


def test_struct_layouts_match_header():
    from b200rag import _ffi
    assert ctypes.sizeof(_ffi.Cand) == 16 and _ffi.CAND_DTYPE.itemsize == 16
    assert ctypes.sizeof(_ffi.Config) == 40
    assert ctypes.sizeof(_ffi.Query) == 24 + 5 * 8
    assert ctypes.sizeof(_ffi.Stats) == 56


def test_compute_fails_loudly_without_gpu(built_lib):
    from b200rag import _ffi
    if _ffi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_ffi.B200RagError) as ei:
        _ffi.Shard(dim=1024)
    assert ei.value.code == _ffi.ERR_NOGPU and "no CPU fallback" in str(ei.value)
    from b200rag.compat import RetrievalConfig, RetrievalError, EmbeddingResult
    from b200rag.retriever import B200Retriever
    r = B200Retriever(RetrievalConfig(), embedding_dim=256)
    with pytest.raises(RetrievalError) as e2:
        r.search(EmbeddingResult(dense=[1.0] * 256))
    assert "connect" in str(e2.value).lower()     # API layer maps this to 503 (api/v1/query.py:151-161)


def test_bad_config_is_rejected(built_lib):
    from b200rag import _ffi
    for kw in ({"dim": 100}, {"dim": 2048}, {"vocab": 0}, {"docs_per_block": 1000}):
        with pytest.raises(_ffi.B200RagError) as ei:
            _ffi.Shard(**{"dim": 256, **kw})
        assert ei.value.code == _ffi.ERR_INVALID
