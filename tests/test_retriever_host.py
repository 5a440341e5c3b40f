"""Host logic of B200Retriever (no GPU): the same requests as the reference's own QdrantRetriever builds, the same
results as the reference's plugin code produces when its qdrant_client is the oracle-backed test double.
Skipped where /root/reference is absent (GPU box); the golden fixture carries the expectations there."""
import json
import os

import numpy as np
import pytest

import conftest
from data_small import DIM, make_chunks, make_queries, result_rows
from oracle_shard import OracleShard

pytestmark = pytest.mark.skipif(not conftest.HAVE_REFERENCE, reason="reference checkout not present")


def _pair(**cfg):
    from audio_rag.config import RetrievalConfig
    from audio_rag.retrieval import RetrievalRegistry
    from b200rag.retriever import B200Retriever
    conf = RetrievalConfig(qdrant_in_memory=True, **cfg)
    ref = RetrievalRegistry.create("qdrant", config=conf, embedding_dim=DIM)
    ours = RetrievalRegistry.create("b200", config=conf, embedding_dim=DIM)
    assert isinstance(ours, B200Retriever)
    ours._shard = OracleShard(dim=DIM)
    return ref, ours


def _types():
    from audio_rag.core import AudioChunk, EmbeddingResult, SparseVector
    return AudioChunk, EmbeddingResult, SparseVector


def test_registry_and_types():
    from audio_rag.core import BaseRetriever
    from audio_rag.retrieval import RetrievalRegistry
    from b200rag import compat
    from b200rag.retriever import B200Retriever
    assert compat.HAVE_REFERENCE and issubclass(B200Retriever, BaseRetriever)
    assert "b200" in RetrievalRegistry and RetrievalRegistry.get("b200") is B200Retriever


@pytest.mark.parametrize("coll_sparse", [True, False])
@pytest.mark.parametrize("query_sparse", [True, False])
@pytest.mark.parametrize("search_type", ["dense", "sparse", "hybrid", None])
@pytest.mark.parametrize("threshold", [0.0, 0.3])
def test_plan_matches_reference_request(coll_sparse, query_sparse, search_type, threshold):
    """Branch/fallback/limits (R8, R11): compare our plan with the request the reference sends to qdrant."""
    A, E, S = _types()
    ref, ours = _pair(score_threshold=threshold, top_k=7)
    chunks, embs = make_chunks(30, 1, "c", A, E, S, sparse=coll_sparse)
    ref.add(chunks, embs, "col")
    ours.add(chunks, embs, "col")
    q = make_queries(1, 5, 30, 1, E, S, sparse=query_sparse)[0]
    for top_k in (None, 3):
        for flt in (None, {"lang": "en"}):
            ref._get_client().requests.clear()
            ref.search(q, top_k=top_k, collection_name="col", filter_metadata=flt, search_type=search_type)
            req = ref._get_client().requests[-1]
            plan = ours._plan_search(q, top_k, "col", flt, search_type)
            assert plan["top_k"] == req["limit"]
            assert plan["mode"] == {"fusion": "hybrid", "sparse": "sparse", "dense": "dense"}[req["query_kind"]]
            if req["fusion"]:
                assert [l for _, l in req["prefetch"]] == [plan["leg_limit"]] * 2 == [2 * plan["top_k"]] * 2
                assert [u for u, _ in req["prefetch"]] == ["dense", "sparse"]
            assert plan["score_threshold"] == req["score_threshold"]
            assert (plan["filter"] or None) == (None if req["filter"] is None else
                                                {k[len("metadata."):]: v for k, v in req["filter"].items()})


def _compare_all(ref, ours, queries, names, **kw):
    for q in queries:
        for name in names:
            a = result_rows(ref.search(q, collection_name=name, **kw))
            b = result_rows(ours.search(q, collection_name=name, **kw))
            assert a == b, f"{name} {kw}\n ref {a[:3]}\n our {b[:3]}"


def test_results_match_reference_plugin_multi_collection():
    A, E, S = _types()
    ref, ours = _pair(top_k=5)
    data = {"tenant_a": make_chunks(120, 11, "A", A, E, S), "tenant_b": make_chunks(80, 12, "B", A, E, S),
            "legacy": make_chunks(60, 13, "L", A, E, S, sparse=False)}
    # interleaved adds: the global row space mixes the tenants
    for part in range(2):
        for name, (ch, em) in data.items():
            h = len(ch) // 2
            sl = slice(0, h) if part == 0 else slice(h, None)
            ref.add(ch[sl], em[sl], name)
            ours.add(ch[sl], em[sl], name)
    for name in data:
        assert ref.count(name) == ours.count(name) == len(data[name][0])
        assert ref.collection_exists(name) and ours.collection_exists(name)
        assert ref.is_hybrid_collection(name) == ours.is_hybrid_collection(name)
    qs = make_queries(4, 21, 120, 11, E, S) + make_queries(2, 22, 60, 13, E, S, sparse=False)
    for st in ("dense", "sparse", "hybrid"):
        _compare_all(ref, ours, qs, list(data), search_type=st)
        _compare_all(ref, ours, qs[:2], list(data), search_type=st, top_k=17)
    for flt in ({"lang": "de"}, {"source": "A-1.wav", "lang": "en"}, {"tags": "a"}, {"missing": 1}, {"idx": 4}):
        _compare_all(ref, ours, qs[:3], ["tenant_a", "legacy"], search_type="hybrid", filter_metadata=flt)
    # unknown collection: created empty, returns [] (R11), then counts 0
    assert ref.search(qs[0], collection_name="nope") == ours.search(qs[0], collection_name="nope") == []
    assert ref.count("nope") == ours.count("nope") == 0 and ours.collection_exists("nope")
    # delete one tenant: the others are unaffected, the name can be reused
    ref.delete_collection("tenant_a")
    ours.delete_collection("tenant_a")
    assert not ours.collection_exists("tenant_a")
    _compare_all(ref, ours, qs[:2], ["tenant_b", "legacy"], search_type="hybrid")
    ch, em = make_chunks(25, 14, "A2", A, E, S)
    ref.add(ch, em, "tenant_a")
    ours.add(ch, em, "tenant_a")
    _compare_all(ref, ours, qs[:2], ["tenant_a", "tenant_b"], search_type="hybrid")
    assert ref.count("tenant_a") == ours.count("tenant_a") == 25


def test_default_collection_threshold_and_batch():
    A, E, S = _types()
    ref, ours = _pair(score_threshold=0.2, top_k=4, search_type="dense", collection_name="dflt")
    ch, em = make_chunks(90, 31, "D", A, E, S, sparse=False)
    ref.add(ch, em)
    ours.add(ch, em)
    qs = make_queries(5, 32, 90, 31, E, S)
    _compare_all(ref, ours, qs, [None])                   # legacy collection: threshold applies (qdrant.py:331)
    assert ref.count() == ours.count() == 90
    batch = ours.search_batch(qs, top_k=4)
    assert [result_rows(b) for b in batch] == [result_rows(ref.search(q)) for q in qs]
    # add() edge cases (qdrant.py:154-160)
    from audio_rag.core import RetrievalError
    ours.add([], [])
    with pytest.raises(RetrievalError):
        ours.add(ch[:2], em[:1])
    with pytest.raises(RetrievalError):
        ours.add(ch[:1], [E(dense=[0.0] * 3)])
    with pytest.raises(RetrievalError):
        ours.add(ch[:1], [E(dense=em[0].dense, sparse=S(indices=[1, 1], values=[1.0, 2.0]))], "h2")
    # results are copies: mutating one does not leak into the store (reranker builds on result.chunk)
    r1 = ours.search(qs[0])
    r1[0].chunk.metadata["poison"] = True
    assert "poison" not in ours.search(qs[0])[0].chunk.metadata


def test_search_batch_arrays_matches_search_batch():
    """The reranker hand-off (arrays + texts, no per-hit objects) returns exactly the hits of search_batch."""
    A, E, S = _types()
    _, ours = _pair(top_k=7)
    for name, seed in (("t1", 61), ("t2", 62)):
        ch, em = make_chunks(70, seed, name.upper(), A, E, S)
        ours.add(ch, em, name)
    qs = make_queries(5, 63, 70, 61, E, S)
    names = ["t1", "t2", "t1", "nope", "t2"]
    objs = ours.search_batch(qs, top_k=7, collection_name=names, search_type="hybrid")
    arr = ours.search_batch_arrays(qs, top_k=7, collection_name=names, search_type="hybrid")
    assert arr["ids"].shape == (5, 7) and arr["sources"] == names
    for b, hits in enumerate(objs):
        assert arr["counts"][b] == len(hits)
        assert arr["texts"][b] == [h.chunk.text for h in hits]
        assert [float(x) for x in arr["scores"][b, :len(hits)]] == [h.score for h in hits]
        if hits:
            m = arr["materialise"](b, 0)
            assert m.chunk.text == hits[0].chunk.text and m.score == hits[0].score and m.source == names[b]
    assert arr["counts"][3] == 0 and (arr["ids"][3] == -1).all()


def test_row_bookkeeping_grows_and_tombstones():
    """The per-row bookkeeping (collection id, tombstone) lives in growable arrays: many small adds beyond the initial
    capacity, interleaved tenants, delete + re-add, and the drop of the storage once nothing is live."""
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever
    A, E, S = _types()
    r = B200Retriever(RetrievalConfig(top_k=3), embedding_dim=DIM)
    r._shard = OracleShard(dim=DIM)
    ch, em = make_chunks(150, 71, "G", A, E, S, sparse=False)
    total = {"g0": 0, "g1": 0, "g2": 0}
    for rep in range(24):                                   # 3 600 rows in 24 adds: crosses the 1024 and 2048 capacities
        name = f"g{rep % 3}"
        r.add(ch, em, name)
        total[name] += 150
    assert len(r._row_coll) == len(r._alive) == len(r._payloads) == 3600 and r._alive.all()
    assert [r.count(n) for n in total] == [1200, 1200, 1200]
    assert (r._row_coll[:450] == np.repeat([0, 1, 2], 150)).all()
    e1 = r._eligible("g1", None)
    assert e1.sum() == 1200 and e1[150:300].all() and not e1[:150].any()
    q = make_queries(1, 72, 150, 71, E, S, sparse=False)[0]
    hits = r.search(q, collection_name="g1", top_k=3)
    assert len(hits) == 3 and all(h.source == "g1" for h in hits)
    r.delete_collection("g1")
    assert r.count("g1") == 0 and r._alive.sum() == 2400 and not r._alive[150:300].any()
    assert r.search(q, collection_name="g1", top_k=3) == []
    r.add(ch[:10], em[:10], "g1")                            # the name is reusable; old rows stay tombstoned
    assert r.count("g1") == 10 and r._eligible("g1", None).sum() == 10 and len(r._row_coll) == 3610
    for n in ("g0", "g1", "g2"):
        r.delete_collection(n)
    assert len(r._row_coll) == len(r._alive) == len(r._payloads) == 0 and r._shard.count == 0
    r.add(ch[:5], em[:5], "g0")
    assert r.count("g0") == 5 and len(r.search(q, collection_name="g0", top_k=3)) == 3


def test_payload_index_equals_per_row_matching():
    """filter_metadata goes through a per-key payload index (value -> rows); it must select exactly the rows the
    per-payload rule `_match` (FieldCondition + MatchValue, qdrant.py:264-268) selects -- list-valued fields, mixed
    numeric types, None, unhashable stored values and unhashable filter values included -- also after further adds."""
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever, _match
    A, E, S = _types()
    r = B200Retriever(RetrievalConfig(top_k=3), embedding_dim=DIM)
    r._shard = OracleShard(dim=DIM)
    ch, em = make_chunks(60, 81, "P", A, E, S, sparse=False)
    odd = [1, 1.0, True, 0, False, None, "1", ["a", "b"], ("a",), [], {"x": 1}, [1, {"y": 2}], "a", 2.5]
    for i, c in enumerate(ch):
        c.metadata = dict(c.metadata)
        if i % 5 != 4:
            c.metadata["odd"] = odd[i % len(odd)]
    probes = [1, True, 0, None, "1", "a", "b", 2.5, {"x": 1}, {"y": 2}, ["a", "b"], "zzz"]

    def brute(key, value):
        return [i for i, p in enumerate(r._payloads) if _match(p.get("metadata"), key, value)]

    r.add(ch[:35], em[:35], "p")
    for v in probes:
        assert sorted(set(r._meta_rows("odd", v).tolist())) == brute("odd", v), v
    r.add(ch[35:], em[35:], "p")                          # the index is extended, not rebuilt
    for v in probes:
        assert sorted(set(r._meta_rows("odd", v).tolist())) == brute("odd", v), v
        elig = r._eligible("p", {"odd": v, "lang": "en"})
        want = [i for i in brute("odd", v) if i in set(brute("lang", "en"))]
        assert np.flatnonzero(elig).tolist() == want, v
    assert r._meta_rows("absent", 1).size == 0


def test_random_operation_sequences_match_reference_plugin():
    """Seeded random sequences of add / delete_collection / search over three tenants, with enough distinct
    (collection, filter) pairs to push masks out of the 64-entry cache and back: after every operation the reference's
    own plugin (over the qdrant_client double) and B200Retriever must agree on results, counts and existence."""
    import random
    A, E, S = _types()
    for seed in (1, 2):
        rng = random.Random(seed)
        ref, ours = _pair(top_k=4)
        names = ["ta", "tb", "legacy"]
        pool = {n: make_chunks(90, 40 + i, n.upper(), A, E, S, sparse=(n != "legacy")) for i, n in enumerate(names)}
        cursor = {n: 0 for n in names}
        qs = make_queries(6, 50 + seed, 90, 40, E, S) + make_queries(2, 60 + seed, 90, 42, E, S, sparse=False)
        filters = [None] + [{"idx": i} for i in range(80)] + [{"lang": "en"}, {"lang": "de"}, {"tags": "a"},
                                                              {"tags": "c", "lang": "de"}, {"source": "TA-1.wav"}]
        for step in range(330):
            op = rng.random()
            name = rng.choice(names)
            if op < 0.15 and cursor[name] < 90:
                n = min(rng.choice([1, 7, 20]), 90 - cursor[name])
                ch, em = pool[name]
                sl = slice(cursor[name], cursor[name] + n)
                ref.add(ch[sl], em[sl], name)
                ours.add(ch[sl], em[sl], name)
                cursor[name] += n
            elif op < 0.17:
                ref.delete_collection(name)
                ours.delete_collection(name)
                cursor[name] = 0                       # the same chunks may be added again (new rows, new ids)
            else:
                q = rng.choice(qs)
                kw = {"search_type": rng.choice(["dense", "sparse", "hybrid", None]), "top_k": rng.choice([None, 1, 9]),
                      "filter_metadata": rng.choice(filters) if rng.random() < 0.8 else None}
                a = result_rows(ref.search(q, collection_name=name, **kw))
                b = result_rows(ours.search(q, collection_name=name, **kw))
                assert a == b, (seed, step, name, kw, a[:2], b[:2])
            for n in names:
                assert ref.count(n) == ours.count(n)
        assert len(ours._masks) <= 64 < ours._next_mask          # the cache did overflow and evict


def test_save_load_host_logic(tmp_path):
    """B200Retriever.save/load (payload chunks, rows.npz, manifest.json, shard files) on the oracle-backed double: a restored
    retriever answers like the original one and like the reference plugin, keeps tombstones and schemas, refuses
    mismatching or corrupt directories."""
    from audio_rag.core import RetrievalError
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever
    A, E, S = _types()
    ref, ours = _pair(top_k=5)
    sets = {"tenant_a": make_chunks(80, 11, "A", A, E, S), "tenant_b": make_chunks(60, 12, "B", A, E, S),
            "legacy": make_chunks(40, 13, "L", A, E, S, sparse=False), "dropped": make_chunks(30, 15, "X", A, E, S)}
    for name, (ch, em) in sets.items():
        ref.add(ch, em, name)
        ours.add(ch, em, name)
    ref.delete_collection("dropped")
    ours.delete_collection("dropped")
    d = str(tmp_path / "saved")
    ours.save(d)
    manifest = json.load(open(os.path.join(d, "manifest.json")))
    rows = np.load(os.path.join(d, "rows.npz"))
    assert manifest["rows"] == 210 and int(rows["alive"].sum()) == 180 and len(rows["row_collection"]) == 210
    assert manifest["collections"]["legacy"]["hybrid"] is False and manifest["collections"]["dropped"]["exists"] is False
    assert len(json.load(open(os.path.join(d, "payloads-0.json"), encoding="utf-8"))) == 210
    assert not [f for f in os.listdir(d) if ".tmp-" in f], "every file is written under a temporary name and renamed"

    def fresh(dim=DIM):
        r = B200Retriever(RetrievalConfig(top_k=5), embedding_dim=dim)
        r._shard = OracleShard(dim=dim)
        return r

    back = fresh()
    assert back.count("probe_only") == 0 and back.is_hybrid_collection("tenant_b") is False   # health probes register names
    back.load(d)
    assert not back.collection_exists("probe_only") and back.is_hybrid_collection("tenant_b") is True, \
        "load() replaces the registry: names registered before it must not alias saved collection ids"
    assert sorted(back._coll_ids.values()) == list(range(len(back._coll_ids)))
    qs = make_queries(4, 21, 80, 11, E, S)
    _compare_all(ref, back, qs, ["tenant_a", "tenant_b", "legacy"], search_type="hybrid")
    _compare_all(ref, back, qs[:2], ["tenant_a"], search_type="dense", filter_metadata={"lang": "en"})
    assert back.count("tenant_a") == 80 and not back.collection_exists("dropped")
    ch, em = make_chunks(10, 16, "A3", A, E, S)
    ref.add(ch, em, "tenant_a")
    back.add(ch, em, "tenant_a")
    _compare_all(ref, back, qs[:2], ["tenant_a"], search_type="hybrid")
    # refusals: non-empty target, other geometry, damaged manifest / payload file
    with pytest.raises(RetrievalError):
        back.load(d)
    with pytest.raises(RetrievalError):
        fresh(dim=DIM * 2).load(d)
    block = json.load(open(os.path.join(d, "payloads-0.json"), encoding="utf-8"))
    json.dump(block[:-3], open(os.path.join(d, "payloads-0.json"), "w", encoding="utf-8"))
    with pytest.raises(RetrievalError):
        fresh().load(d)


def test_golden_fixture_is_current():
    """tests/golden/retrieval_golden.json was produced by make_golden.py from the reference plugin + test double."""
    import make_golden
    path = os.path.join(os.path.dirname(__file__), "golden", "retrieval_golden.json")
    assert os.path.exists(path), "run python tests/golden/make_golden.py"
    assert json.load(open(path)) == json.loads(json.dumps(make_golden.generate()))


def test_compressed_scan_option_reaches_every_shard(monkeypatch):
    """compressed_scan=True / B200RAG_COMPRESSED_SCAN=1 -> set_compression(True) on each shard when the shards are set;
    an engine that refuses (row width the 8-bit scan does not cover) costs a warning, not the retriever."""
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever
    try:
        conf = RetrievalConfig(qdrant_in_memory=True)
    except TypeError:
        conf = RetrievalConfig()
    r = B200Retriever(conf, embedding_dim=1024, compressed_scan=True)
    shards = [OracleShard(dim=1024), OracleShard(dim=1024)]
    r._set_shards(shards, group=object())
    assert all(("set_compression", True) in sh.calls for sh in shards)
    monkeypatch.setenv("B200RAG_COMPRESSED_SCAN", "1")
    r2 = B200Retriever(conf, embedding_dim=1024)
    one = OracleShard(dim=1024)
    r2._shard = one
    assert ("set_compression", True) in one.calls
    monkeypatch.setenv("B200RAG_COMPRESSED_SCAN", "0")
    r3 = B200Retriever(conf, embedding_dim=1024)
    off = OracleShard(dim=1024)
    r3._shard = off
    assert off.calls == []
    r4 = B200Retriever(conf, embedding_dim=DIM, compressed_scan=True)      # 256-wide rows: refused by the engine
    narrow = OracleShard(dim=DIM)
    r4._shard = narrow
    A, E, S = _types()
    ch, em = make_chunks(20, 5, "N", A, E, S)
    r4.add(ch, em, "n")
    assert len(r4.search(make_queries(1, 6, 20, 5, E, S)[0], collection_name="n")) > 0


def test_query_conversion_fast_paths_equal_numpy(built_lib):
    """The single-query fast paths (struct.pack for the dense list, plain-Python sort for a short sparse vector) must hand
    the engine exactly what the general numpy path builds -- same bf16 bits, same sorted terms and weights -- and must
    reject what it rejects."""
    import types
    from audio_rag.core import RetrievalError
    from b200rag import _ffi
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever, _sorted_sparse
    A, E, S = _types()
    r = B200Retriever(RetrievalConfig(qdrant_in_memory=True), embedding_dim=DIM)
    rng = np.random.default_rng(9)
    for trial in range(20):
        dense64 = (rng.standard_normal(DIM) * 10.0 ** rng.integers(-6, 4)).tolist()       # Python doubles, as an embedder's .tolist()
        if trial == 3:
            dense64 = [int(x) for x in rng.integers(-5, 6, DIM)]                            # ints are numbers too
        if trial == 4:
            dense64 = list(np.asarray(dense64, np.float32))                                 # numpy scalars in a list
        nt = int(rng.integers(0, 30))
        idx = rng.choice(250_002, nt, replace=False).tolist()
        val = rng.random(nt).tolist()
        e = E(dense=dense64, sparse=S(indices=idx, values=val))
        q_bits, ip, tt, ww = r._query_arrays([e], "hybrid")
        assert np.array_equal(q_bits, _ffi.normalize_bf16(np.asarray([dense64], dtype=np.float32)))
        o = np.argsort(np.asarray(idx, np.int64), kind="stable")
        assert ip.tolist() == [0, nt] and tt.dtype == np.uint32 and ww.dtype == np.float32
        assert np.array_equal(tt, np.asarray(idx, np.uint32)[o]) and np.array_equal(ww, np.asarray(val, np.float32)[o])
        # ... and the batch path (numpy all the way) agrees with itself on the same query
        qb2, ip2, tt2, ww2 = r._query_arrays([e, e], "hybrid")
        assert np.array_equal(qb2[0], q_bits[0]) and np.array_equal(tt2[:nt], tt) and np.array_equal(ww2[:nt], ww)
    long_idx = rng.choice(250_002, 200, replace=False).tolist()                             # > 64 terms: numpy path
    t_long, _ = _sorted_sparse(S(indices=long_idx, values=[1.0] * 200), 250_002)
    assert np.array_equal(t_long, np.sort(np.asarray(long_idx, np.uint32)))
    for bad in (S(indices=[5, 5], values=[1.0, 2.0]), S(indices=[-1], values=[1.0]), S(indices=[250_002], values=[1.0]),
                S(indices=[1, 2], values=[1.0])):
        with pytest.raises(RetrievalError):
            _sorted_sparse(bad, 250_002)
    with pytest.raises(RetrievalError):
        r._query_arrays([E(dense=[0.5] * (DIM - 1), sparse=None)], "dense")                 # wrong width
    with pytest.raises(Exception):
        r._query_arrays([E(dense=["x"] * DIM, sparse=None)], "dense")                       # not numbers
