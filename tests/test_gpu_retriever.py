"""B200Retriever on a real B200: against the committed golden fixture (expected outputs of the reference's own
plugin code, tests/make_golden.py) and against its oracle-backed twin through the same adapter code."""
import json
import os

import numpy as np
import pytest

from data_small import DIM, make_chunks, make_queries, result_rows
from oracle_shard import OracleShard

pytestmark = pytest.mark.gpu


def _types():
    from b200rag.compat import AudioChunk, EmbeddingResult, SparseVector
    return AudioChunk, EmbeddingResult, SparseVector


def _retriever(gpu, **cfg):
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever
    try:
        conf = RetrievalConfig(qdrant_in_memory=True, **cfg)
    except TypeError:
        conf = RetrievalConfig(**cfg)
    return B200Retriever(conf, embedding_dim=DIM, device=gpu, docs_per_block=1024)


def test_golden_fixture(gpu):
    import make_golden
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "retrieval_golden.json")))
    assert gold["dim"] == DIM
    r = _retriever(gpu)
    make_golden.load(r, _types())
    got = make_golden.run_cases(r, _types())
    assert len(got) == len(gold["cases"])
    for g, e in zip(got, gold["cases"]):
        assert g == e, f"case {e['search_type']} k={e['top_k']} {e['collection']} {e['filter']} q{e['query']}"
    assert r._shard is not None and type(r._shard).__name__ == "Shard"
    r.close()


def test_adapter_gpu_vs_oracle_twin(gpu):
    A, E, S = _types()
    real, twin = _retriever(gpu, top_k=6), _retriever(gpu, top_k=6)
    twin._shard = OracleShard(dim=DIM)
    data = {"t1": make_chunks(700, 41, "T1", A, E, S), "t2": make_chunks(300, 42, "T2", A, E, S),
            "old": make_chunks(200, 43, "O", A, E, S, sparse=False)}
    for name, (ch, em) in data.items():
        for s in range(0, len(ch), 170):
            real.add(ch[s:s + 170], em[s:s + 170], name)
            twin.add(ch[s:s + 170], em[s:s + 170], name)
    qs = make_queries(6, 51, 700, 41, E, S)
    for st in ("dense", "sparse", "hybrid"):
        for name in list(data) + ["unknown"]:
            for flt in (None, {"lang": "en"}):
                for q in qs[:3]:
                    a = result_rows(real.search(q, collection_name=name, search_type=st, filter_metadata=flt))
                    b = result_rows(twin.search(q, collection_name=name, search_type=st, filter_metadata=flt))
                    assert a == b, (st, name, flt)
    # batched, mixed tenants in one call (config-4 style: one pass, one mask per query)
    names = ["t1", "t2", "t1", "old", "t2", "t1"]
    a = real.search_batch(qs, top_k=10, collection_name=names, search_type="hybrid")
    b = twin.search_batch(qs, top_k=10, collection_name=names, search_type="hybrid")
    assert [result_rows(x) for x in a] == [result_rows(x) for x in b]
    # delete + re-add keeps parity
    real.delete_collection("t1"); twin.delete_collection("t1")
    ch, em = make_chunks(50, 44, "T1b", A, E, S)
    real.add(ch, em, "t1"); twin.add(ch, em, "t1")
    for q in qs[:2]:
        for name in ("t1", "t2"):
            assert result_rows(real.search(q, collection_name=name)) == result_rows(twin.search(q, collection_name=name))
    assert real.count("t1") == 50 and real.count("t2") == 300
    for name in ("t1", "t2", "old"):
        real.delete_collection(name)
    assert real._shard.count == 0
    real.close()


def test_save_load_round_trip(gpu, tmp_path):
    """Persistence: a saved retriever restored into a fresh one answers every search identically
    (ids, scores, payloads), keeps collection schemas, tombstones and counts, and accepts further adds."""
    A, E, S = _types()
    r = _retriever(gpu, top_k=6)
    data = {"t1": make_chunks(600, 41, "T1", A, E, S), "t2": make_chunks(250, 42, "T2", A, E, S),
            "old": make_chunks(120, 43, "O", A, E, S, sparse=False), "gone": make_chunks(90, 45, "G", A, E, S)}
    for name, (ch, em) in data.items():
        r.add(ch, em, name)
    r.delete_collection("gone")
    qs = make_queries(4, 51, 600, 41, E, S)

    def snapshot(x):
        out = []
        for st in ("dense", "sparse", "hybrid"):
            for name in ("t1", "t2", "old", "gone"):
                for flt in (None, {"lang": "en"}):
                    for q in qs:
                        out.append(result_rows(x.search(q, collection_name=name, search_type=st, filter_metadata=flt)))
        return out

    before = snapshot(r)
    d = str(tmp_path / "idx")
    r.save(d)
    assert sorted(os.listdir(d)) == ["manifest.json", "payloads-0.json", "rows.npz", "shard-0.bin"]
    r2 = _retriever(gpu, top_k=6)
    r2.load(d)
    assert snapshot(r2) == before
    assert [r2.count(n) for n in ("t1", "t2", "old")] == [600, 250, 120]
    assert r2.is_hybrid_collection("t1") and not r2.is_hybrid_collection("old")
    # (searching a deleted name re-creates it empty, like the reference: qdrant.py:248) -> same state on both sides
    assert r2.collection_exists("gone") == r.collection_exists("gone") and r2.count("gone") == 0
    ch, em = make_chunks(40, 46, "T2b", A, E, S)
    r.add(ch, em, "t2"); r2.add(ch, em, "t2")
    assert snapshot(r2) == snapshot(r)
    # a foreign file is refused
    from b200rag.compat import RetrievalError
    open(os.path.join(d, "shard-0.bin"), "wb").write(b"not a shard file at all")
    r3 = _retriever(gpu, top_k=6)
    with pytest.raises(RetrievalError):
        r3.load(d)
    # ... and leaves a clean, usable retriever behind (a failed load must not corrupt later adds)
    r3.add(ch, em, "fresh")
    assert r3.count("fresh") == 40 and len(r3.search(qs[0], collection_name="fresh", search_type="hybrid")) == 6
    for x in (r, r2, r3):
        x.close()


def test_compressed_scan_option_changes_nothing_but_the_scan(gpu):
    """B200Retriever(compressed_scan=True) (or B200RAG_COMPRESSED_SCAN=1): the opt-in 8-bit candidate scan behind the
    plugin -- same hits, same scores, and the engine says it took the 8-bit scan (1024-wide rows, BGE-M3's width)."""
    A, E, S = _types()
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever
    try:
        conf = RetrievalConfig(qdrant_in_memory=True, top_k=6)
    except TypeError:
        conf = RetrievalConfig(top_k=6)
    plain = B200Retriever(conf, embedding_dim=1024, device=gpu, docs_per_block=1024)
    comp = B200Retriever(conf, embedding_dim=1024, device=gpu, docs_per_block=1024, compressed_scan=True)
    ch, em = make_chunks(900, 61, "C", A, E, S, dim=1024)
    for s in range(0, 900, 300):
        plain.add(ch[s:s + 300], em[s:s + 300], "c1")
        comp.add(ch[s:s + 300], em[s:s + 300], "c1")
    for q in make_queries(4, 62, 900, 61, E, S, dim=1024):
        for st in ("dense", "hybrid"):
            a = result_rows(plain.search(q, collection_name="c1", search_type=st))
            b = result_rows(comp.search(q, collection_name="c1", search_type=st))
            assert a == b and len(a) > 0
            assert comp._shard.stats()["dense_path"] == 3 and plain._shard.stats()["dense_path"] == 1
    # a width the 8-bit scan does not cover: the option is dropped with a warning, the retriever works
    small = B200Retriever(conf, embedding_dim=DIM, device=gpu, docs_per_block=1024, compressed_scan=True)
    ch2, em2 = make_chunks(200, 63, "D", A, E, S)
    small.add(ch2, em2, "c2")
    assert len(small.search(make_queries(1, 64, 200, 63, E, S)[0], collection_name="c2", search_type="hybrid")) > 0
    plain.close()
    comp.close()
    small.close()
