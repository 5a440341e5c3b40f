"""The reference's own orchestration (AudioRAG / QueryPipeline, unchanged) running on top of B200Retriever, injected
through the seam AudioRAG itself uses (pipeline/orchestrator.py:64-65,73-74).
CPU: the retriever's shard is the oracle double (host logic under test).  The GPU twin of this test is
tests/test_gpu_retriever.py (the reference checkout does not travel to the GPU box)."""
import numpy as np
import pytest

import conftest
from data_small import DIM, make_chunks, make_queries
from oracle_shard import OracleGroup, OracleShard

pytestmark = pytest.mark.skipif(not conftest.HAVE_REFERENCE, reason="reference checkout not present")


class StubEmbedder:
    """What QueryPipeline needs from an embedder (pipeline/query.py:62-67,137-138)."""
    is_loaded = True
    dimension = DIM
    vram_required = 0.0
    supports_sparse = True

    def __init__(self, queries):
        self.queries = queries
        self.calls = []

    def load(self):
        pass

    def unload(self):
        pass

    def embed_query(self, text):
        self.calls.append(text)
        return self.queries[int(text.split("#")[1])]


def _rag(tmp_path):
    from audio_rag.config import AudioRAGConfig
    from audio_rag.pipeline import AudioRAG
    cfg = AudioRAGConfig(reranking={"backend": "none"}, generation={"backend": "none"},
                         data_dir=str(tmp_path / "d"), cache_dir=str(tmp_path / "c"))
    return AudioRAG(cfg), cfg


@pytest.mark.parametrize("n_shards", [1, 2])
def test_audio_rag_query_runs_on_b200_retriever(tmp_path, n_shards):
    """n_shards = 2: the reference's AudioRAG.query() on a retriever that owns TWO shards."""
    from audio_rag.core import AudioChunk, EmbeddingResult, SparseVector
    from b200rag.retriever import B200Retriever
    rag, cfg = _rag(tmp_path)
    retr = B200Retriever(cfg.retrieval, embedding_dim=DIM)
    shards = [OracleShard(dim=DIM) for _ in range(n_shards)]
    retr._set_shards(shards, OracleGroup(shards) if n_shards > 1 else None)
    ch, em = make_chunks(150, 61, "P", AudioChunk, EmbeddingResult, SparseVector)
    for s in range(0, 150, 50):
        retr.add(ch[s:s + 50], em[s:s + 50], "tenant_a")
    assert all(sh.count > 0 for sh in shards)
    qs = make_queries(3, 62, 150, 61, EmbeddingResult, SparseVector)
    rag._embedder = StubEmbedder(qs)
    rag._retriever = retr                       # before first use: propagated to both pipelines
    res = rag.query("question #0", collection_name="tenant_a", top_k=10, generate_answer=False)
    direct = retr.search(qs[0], top_k=10, collection_name="tenant_a", search_type="hybrid")
    got = getattr(res, "results", None) or getattr(res, "chunks", None) or res["results"]
    assert [r.chunk.text for r in got] == [r.chunk.text for r in direct] and len(direct) == 10
    assert retr._shard.calls[-1][0] == "hybrid"
    # get_context passes no search_type -> config default (qdrant.py:250); status() -> count()
    ctx = rag.get_context("question #1", collection_name="tenant_a", top_k=3)
    assert isinstance(ctx, str) and direct is not None
    assert retr.count("tenant_a") == 150
    st = rag.status()
    assert isinstance(st, dict)
