"""Types crossing the retrieval boundary.

When the reference package is importable (``audio_rag`` on sys.path) its own classes are used, so a
``B200Retriever`` is a real ``audio_rag.core.BaseRetriever`` and returns real ``RetrievalResult`` objects.
Otherwise (e.g. on a GPU box without the reference checkout) structurally identical stand-ins are defined:
same field names and defaults as /root/reference/src/audio_rag/core/base.py:29-61 and
src/audio_rag/config/schema.py:58-69, nothing more.
"""
from __future__ import annotations

import functools
import logging
import time
from abc import ABC, abstractmethod
from dataclasses import dataclass

HAVE_REFERENCE = False
try:  # pragma: no cover - depends on the environment
    from audio_rag.core import (AudioChunk, BaseRetriever, EmbeddingResult, RetrievalError,  # type: ignore
                                RetrievalResult, SparseVector)
    from audio_rag.config import RetrievalConfig  # type: ignore
    HAVE_REFERENCE = True
except Exception:  # ImportError or a missing transitive dependency of the reference package
    @dataclass
    class AudioChunk:  # core/base.py:29-36
        text: str
        start: float
        end: float
        speaker: str | None = None
        metadata: dict | None = None

    @dataclass
    class SparseVector:  # core/base.py:39-46
        indices: list[int]
        values: list[float]

        def to_dict(self) -> dict[int, float]:
            return dict(zip(self.indices, self.values))

    @dataclass
    class EmbeddingResult:  # core/base.py:49-53
        dense: list[float]
        sparse: SparseVector | None = None

    @dataclass
    class RetrievalResult:  # core/base.py:56-61
        chunk: AudioChunk
        score: float
        source: str | None = None

    class RetrievalError(Exception):  # core/exceptions.py:44-46
        pass

    class BaseRetriever(ABC):  # core/base.py:170-190
        @abstractmethod
        def add(self, chunks, embeddings, collection_name=None) -> None: ...

        @abstractmethod
        def search(self, query_embedding, top_k=None, collection_name=None, filter_metadata=None): ...

    @dataclass
    class RetrievalConfig:  # config/schema.py:58-69 (validation ranges are enforced by the reference's pydantic model)
        backend: str = "qdrant"
        collection_name: str = "audio_rag"
        search_type: str = "hybrid"
        top_k: int = 5
        score_threshold: float = 0.0
        qdrant_host: str = "localhost"
        qdrant_port: int = 6333
        qdrant_in_memory: bool = False
        dense_weight: float = 0.7
        sparse_weight: float = 0.3


def get_logger(name: str) -> logging.Logger:
    return logging.getLogger(name)


def timed(fn):
    """Wall-time logging like the reference's utils/decorators.py:14-23 (@timed on search/add, qdrant.py:140,227)."""
    log = logging.getLogger(fn.__module__)

    @functools.wraps(fn)
    def wrapper(*a, **kw):
        t0 = time.perf_counter()
        try:
            return fn(*a, **kw)
        finally:
            log.info("%s took %.3fs", fn.__name__, time.perf_counter() - t0)
    return wrapper
