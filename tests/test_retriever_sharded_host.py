"""Host logic of the SHARDED plugin (no GPU): B200Retriever owning several shards must answer, bit for bit, like a
B200Retriever owning one shard -- routing of add() batches, global row ids, per-shard mask slices, compaction after
delete_collection, persistence.  The shards are oracle doubles (tests/oracle_shard.py); the GPU twin of this file is
tests/test_gpu_group.py."""
import numpy as np
import pytest

from data_small import DIM, make_chunks, make_queries, result_rows
from oracle_shard import OracleGroup, OracleShard


def _types():
    from b200rag.compat import AudioChunk, EmbeddingResult, SparseVector
    return AudioChunk, EmbeddingResult, SparseVector


def _retr(n_shards, **kw):
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever
    try:
        conf = RetrievalConfig(qdrant_in_memory=True, top_k=6)
    except TypeError:
        conf = RetrievalConfig(top_k=6)
    r = B200Retriever(conf, embedding_dim=DIM, **kw)
    shards = [OracleShard(dim=DIM) for _ in range(n_shards)]
    r._set_shards(shards, OracleGroup(shards) if n_shards > 1 else None)
    return r


def _same(a, b, qs, names, **kw):
    for name in names:
        for q in qs:
            ra = result_rows(a.search(q, collection_name=name, **kw))
            rb = result_rows(b.search(q, collection_name=name, **kw))
            assert ra == rb, (name, kw)


def test_sharded_plugin_equals_single_shard(tmp_path):
    A, E, S = _types()
    one, three = _retr(1, compact_dead_fraction=0.3), _retr(3, compact_dead_fraction=0.3)
    data = {"t1": make_chunks(500, 41, "T1", A, E, S), "t2": make_chunks(260, 42, "T2", A, E, S),
            "old": make_chunks(120, 43, "O", A, E, S, sparse=False)}
    for name, (ch, em) in data.items():
        for s in range(0, len(ch), 90):                       # many small adds: routed whole to the emptiest shard
            one.add(ch[s:s + 90], em[s:s + 90], name)
            three.add(ch[s:s + 90], em[s:s + 90], name)
    loads = [g.n for g in three._shard_rows]
    assert sum(loads) == 880 and max(loads) - min(loads) <= 90, loads
    # ids are insertion order whatever the shard
    for g in three._shard_rows:
        assert (np.diff(g.view) > 0).all()
    assert sorted(np.concatenate([g.view for g in three._shard_rows]).tolist()) == list(range(880))
    qs = make_queries(5, 51, 500, 41, E, S)
    for st in ("dense", "sparse", "hybrid"):
        _same(one, three, qs[:3], ["t1", "t2", "old", "unknown"], search_type=st)
        _same(one, three, qs[:2], ["t1", "t2"], search_type=st, filter_metadata={"lang": "en"})
    _same(one, three, qs[:2], ["t1"], search_type="hybrid", top_k=100)
    names = ["t1", "t2", "t1", "old", "t2"]
    ba = one.search_batch(qs, top_k=10, collection_name=names, search_type="hybrid")
    bb = three.search_batch(qs, top_k=10, collection_name=names, search_type="hybrid")
    assert [result_rows(x) for x in ba] == [result_rows(x) for x in bb]

    # delete a big tenant: every shard crosses the dead-fraction threshold and compacts; ids never change
    for r in (one, three):
        r.delete_collection("t1")
        assert r._stored == 380 and r.count("t1") == 0 and r.count("t2") == 260
        assert sum(s.count for s in r._shards) == 380
    assert three._payloads[0] is None and three._payloads[500] is not None
    _same(one, three, qs[:3], ["t2", "old", "t1"], search_type="hybrid")
    ch, em = make_chunks(70, 44, "T1b", A, E, S)
    one.add(ch, em, "t1")
    three.add(ch, em, "t1")
    assert three._row_coll.shape[0] == 950
    _same(one, three, qs[:3], ["t1", "t2"], search_type="hybrid")
    _same(one, three, qs[:2], ["t1", "t2"], search_type="dense", filter_metadata={"lang": "de"})

    # persistence of the sharded layout
    d = str(tmp_path / "snap")
    three.save(d)
    back = _retr(3, compact_dead_fraction=0.3)
    back.load(d)
    _same(one, back, qs[:3], ["t1", "t2", "old"], search_type="hybrid")
    with pytest.raises(Exception):
        _retr(2).load(d)                                     # another shard count is refused

    # a small delete stays a tombstone where a shard's dead fraction is below the threshold (the decision is per shard);
    # a forced compaction drops the rest
    for r in (one, back):
        r.delete_collection("old")
        assert r.count("old") == 0 and 330 <= r._stored <= 450
        r._compact(force=True)
        assert r._stored == 330 and sum(s.count for s in r._shards) == 330
    assert one._stored == 330
    _same(one, back, qs[:3], ["t1", "t2", "old"], search_type="hybrid")
    for r in (one, back):
        for name in ("t1", "t2"):
            r.delete_collection(name)
        assert r._stored == 0 and len(r._payloads) == 0 and all(s.count == 0 for s in r._shards)


def test_bulk_add_is_split_over_the_shards():
    A, E, S = _types()
    r = _retr(2, device_add_rows=10 ** 9)
    ch, em = make_chunks(300, 7, "B", A, E, S)
    r._route = lambda n, _orig=r._route: _orig(n) if n < 300 else [(0, 0, 170), (1, 170, 300)]
    r.add(ch, em, "bulk")
    assert [g.n for g in r._shard_rows] == [170, 130]
    assert r._shard_rows[1].view[0] == 170
    # the real router: bulk batches are water-filled, small ones go to the emptiest shard
    r2 = _retr(4)
    r2._shard_rows[0].append(np.arange(5000))
    parts = r2._route(4 * 4096 + 3)
    assert sum(hi - lo for _, lo, hi in parts) == 4 * 4096 + 3 and [lo for _, lo, _ in parts] == sorted(lo for _, lo, _ in parts)
    after = [r2._shard_rows[s].n + sum(hi - lo for ss, lo, hi in parts if ss == s) for s in range(4)]
    assert max(after) - min(after) <= 1, after
    assert r2._route(100) == [(1, 0, 100)]


def test_mask_cache_never_evicts_a_mask_of_the_running_batch():
    """a search_batch with more distinct (collection, filter) pairs than the mask cache holds must not drop
    masks whose ids were already handed to the engine for the same call."""
    A, E, S = _types()
    r = _retr(1)
    n_coll = 80
    for c in range(n_coll):
        ch, em = make_chunks(6, 100 + c, f"C{c}", A, E, S)
        r.add(ch, em, f"c{c}")
    qs = make_queries(n_coll, 9, 6, 100, E, S)
    out = r.search_batch(qs, top_k=3, collection_name=[f"c{c}" for c in range(n_coll)], search_type="hybrid")
    assert all(len(x) == 3 for x in out)
    for c, res in enumerate(out):
        assert all(x.chunk.text.startswith(f"C{c} ") for x in res)
    assert len(r._masks) <= n_coll and r._next_mask >= n_coll
    r.search(qs[0], collection_name="c0")
    assert len(r._masks) <= 64                                 # trimmed once no batch is using the surplus
