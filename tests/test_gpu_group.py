"""GPU tests of the in-process multi-shard path (b200rag_group_* and the sharded B200Retriever): several shards --
on one GPU here, one per GPU on a multi-GPU box (B200RAG_TEST_DEVICES) -- must return, bit for bit, what one shard
holding every row returns, and what the oracle returns."""
import os

import numpy as np
import pytest

from data_small import DIM, make_chunks, make_queries, result_rows
from helpers import Corpus, assert_result_equal, oracle_search
from oracle_shard import OracleGroup, OracleShard

pytestmark = pytest.mark.gpu


def _devices(gpu, n):
    env = os.environ.get("B200RAG_TEST_DEVICES")
    if env:
        devs = [int(x) for x in env.split(",")]
        return [devs[i % len(devs)] for i in range(n)]
    return [gpu] * n


def test_group_search_equals_single_shard_and_oracle(gpu):
    from b200rag import Shard, ShardGroup, normalize_bf16
    from b200rag.synth import pack_mask
    c = Corpus(20_000, dim=1024, vocab=60_013)
    one = Shard(dim=1024, vocab=c.vocab, device=gpu, docs_per_block=2048)
    one.add(c.bits, c.indptr, c.terms, c.w)
    # three shards, rows dealt out in interleaved runs (ids explicit), so every shard holds rows of every id range
    runs = np.array_split(np.arange(c.n), 23)
    owner = [i % 3 for i in range(len(runs))]
    shards = [Shard(dim=1024, vocab=c.vocab, device=d, docs_per_block=2048) for d in _devices(gpu, 3)]
    rows_of = [[], [], []]
    for r, o in zip(runs, owner):
        s, e = int(r[0]), int(r[-1]) + 1
        shards[o].add(c.bits[s:e], c.indptr[s:e + 1] - c.indptr[s], c.terms[c.indptr[s]:c.indptr[e]],
                      c.w[c.indptr[s]:c.indptr[e]], ids=np.arange(s, e, dtype=np.int64))
        rows_of[o].append(np.arange(s, e))
    rows_of = [np.concatenate(x) for x in rows_of]
    grp = ShardGroup(shards)
    rng = np.random.default_rng(1)
    masks = {0: rng.random(c.n) < 0.25, 1: rng.random(c.n) < 0.002}
    for m, bits in masks.items():
        one.mask_set(m, pack_mask(bits), c.n)
        for sh, rows in zip(shards, rows_of):
            sh.mask_set(m, pack_mask(bits[rows]), len(rows))
    qf, ip, tt, ww = c.queries(6)
    qb = normalize_bf16(qf)
    for mode, k, mids in (("dense", 10, None), ("sparse", 10, None), ("hybrid", 10, None), ("hybrid", 100, None),
                          ("dense", 100, None), ("sparse", 100, None),
                          ("hybrid", 10, np.asarray([0, 1, -1, 0, 1, 0], np.int32)),
                          ("dense", 5, np.asarray([1, 1, 1, 1, 1, 1], np.int32))):
        a = one.search(mode, k, qb, ip, tt, ww, mask_ids=mids)
        g = grp.search(mode, k, qb, ip, tt, ww, mask_ids=mids)
        for x, y in zip(a, g):
            assert np.array_equal(x, y), f"group != single shard: {mode} k={k} masks={mids is not None}"
        for q in range(6):
            elig = None if mids is None or mids[q] < 0 else masks[int(mids[q])]
            e_i, e_s = oracle_search(c, mode, qb[q], tt[ip[q]:ip[q + 1]], ww[ip[q]:ip[q + 1]], elig, k)
            assert_result_equal(g[0][q], g[1][q], int(g[2][q]), e_i, e_s, ctx=f"group {mode} k={k} q{q}")
    assert grp.stats()["kernel_launches"] > one.stats()["kernel_launches"]
    # one query at a time (the plugin's search()), and a batch on the tcgen05 path
    for q in range(2):
        a = one.search("hybrid", 10, qb[q:q + 1], ip[q:q + 2] - ip[q], tt[ip[q]:ip[q + 1]], ww[ip[q]:ip[q + 1]])
        g = grp.search("hybrid", 10, qb[q:q + 1], ip[q:q + 2] - ip[q], tt[ip[q]:ip[q + 1]], ww[ip[q]:ip[q + 1]])
        assert all(np.array_equal(x, y) for x, y in zip(a, g))
    # exhaustive legs through the group
    for sh in shards:
        sh.set_exhaustive(True)
    g = grp.search("hybrid", 10, qb, ip, tt, ww)
    for sh in shards:
        sh.set_exhaustive(False)
    a = one.search("hybrid", 10, qb, ip, tt, ww)
    assert all(np.array_equal(x, y) for x, y in zip(a, g))
    grp.close()
    for sh in shards + [one]:
        sh.close()


def _types():
    from b200rag.compat import AudioChunk, EmbeddingResult, SparseVector
    return AudioChunk, EmbeddingResult, SparseVector


def _retriever(devices, **cfg):
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever
    try:
        conf = RetrievalConfig(qdrant_in_memory=True, **cfg)
    except TypeError:
        conf = RetrievalConfig(**cfg)
    return B200Retriever(conf, embedding_dim=DIM, devices=devices, docs_per_block=1024, compact_dead_fraction=0.3,
                         device_add_rows=128)


def test_sharded_retriever_equals_single_shard_and_oracle_twin(gpu, tmp_path):
    """The plugin with three shards (one process) == the plugin with one shard == its oracle-backed twin, through
    add / search / search_batch / filters / delete_collection + compaction / re-add / save + load.  Adds of >= 128 rows
    take the GPU-side normalise path (b200rag_add_f32), smaller ones the host routine."""
    A, E, S = _types()
    one, three, twin = _retriever([gpu], top_k=6), _retriever(_devices(gpu, 3), top_k=6), _retriever([gpu], top_k=6)
    tw = [OracleShard(dim=DIM) for _ in range(3)]
    twin._set_shards(tw, OracleGroup(tw))
    data = {"t1": make_chunks(700, 41, "T1", A, E, S), "t2": make_chunks(300, 42, "T2", A, E, S),
            "old": make_chunks(200, 43, "O", A, E, S, sparse=False)}
    for name, (ch, em) in data.items():
        for s in range(0, len(ch), 170):
            for r in (one, three, twin):
                r.add(ch[s:s + 170], em[s:s + 170], name)
    assert three.n_shards == 3 and all(sh.count > 0 for sh in three._shards)
    assert type(three._group).__name__ == "ShardGroup"
    qs = make_queries(6, 51, 700, 41, E, S)

    def same(names, qq, **kw):
        for name in names:
            for q in qq:
                a = result_rows(one.search(q, collection_name=name, **kw))
                b = result_rows(three.search(q, collection_name=name, **kw))
                t = result_rows(twin.search(q, collection_name=name, **kw))
                assert a == b == t, (name, kw)

    for st in ("dense", "sparse", "hybrid"):
        same(list(data) + ["unknown"], qs[:3], search_type=st)
        same(["t1", "t2"], qs[:2], search_type=st, filter_metadata={"lang": "en"})
    same(["t1"], qs[:2], search_type="hybrid", top_k=100)
    names = ["t1", "t2", "t1", "old", "t2", "t1"]
    got = [[result_rows(x) for x in r.search_batch(qs, top_k=10, collection_name=names, search_type="hybrid")]
           for r in (one, three, twin)]
    assert got[0] == got[1] == got[2]
    arr = three.search_batch_arrays(qs, top_k=10, collection_name=names, search_type="hybrid")
    assert [[t for t in row] for row in arr["texts"]] == [[x[0] for x in res] for res in got[1]]
    assert arr["ids"].shape == (6, 10) and (arr["counts"] == [len(r) for r in got[1]]).all()
    # delete + compaction + re-add
    for r in (one, three, twin):
        r.delete_collection("t1")
        assert r._stored == 500 and sum(s.count for s in r._shards) == 500
    same(["t1", "t2", "old"], qs[:3], search_type="hybrid")
    ch, em = make_chunks(150, 44, "T1b", A, E, S)
    for r in (one, three, twin):
        r.add(ch, em, "t1")
    same(["t1", "t2"], qs[:3], search_type="hybrid")
    same(["t2"], qs[:2], search_type="sparse", filter_metadata={"lang": "de"})
    # persistence of the sharded layout on real shards
    d = str(tmp_path / "snap")
    three.save(d)
    back = _retriever(_devices(gpu, 3), top_k=6)
    back.load(d)
    for name in ("t1", "t2", "old"):
        for q in qs[:2]:
            assert result_rows(back.search(q, collection_name=name)) == result_rows(one.search(q, collection_name=name))
    for r in (one, three, back):
        r.close()


def test_integration_md_snippet_verbatim(gpu):
    """INTEGRATION.md section 4 builds Shard + ShardedSearcher WITHOUT calling set_stream.  The
    searcher must tie the shard to torch's current stream itself, or its buffers' fills, the result read-back and the
    library's kernels run unordered on two streams.  World 1 here (the N > 1 form is tools/dist_check.py); many
    searches back to back on a NON-default torch stream, so that a missing tie shows up as stale results."""
    import torch
    from b200rag import Shard, normalize_bf16
    from b200rag.dist import ShardedSearcher, shard_bounds
    c = Corpus(30_000, dim=1024, vocab=60_013)
    world, rank, local_rank = 1, 0, gpu
    n_rows_total = c.n
    qf, ip, tt, ww = c.queries(12)
    q_bits = normalize_bf16(qf)
    side = torch.cuda.Stream(device=torch.device("cuda", gpu))
    with torch.cuda.stream(side):
        # ---- the snippet
        lo, hi = shard_bounds(n_rows_total, world, rank)
        shard = Shard(dim=1024, vocab=c.vocab, device=local_rank, row_base=lo)
        shard.add(c.bits[lo:hi], c.indptr[lo:hi + 1] - c.indptr[lo], c.terms[c.indptr[lo]:c.indptr[hi]],
                  c.w[c.indptr[lo]:c.indptr[hi]])
        searcher = ShardedSearcher(shard, torch.device("cuda", local_rank))
        for q in range(12):
            sp_indptr, sp_terms, sp_weights = ip[q:q + 2] - ip[q], tt[ip[q]:ip[q + 1]], ww[ip[q]:ip[q + 1]]
            ids, scores, counts = searcher.search("hybrid", 10, q_bits[q:q + 1], sp_indptr, sp_terms, sp_weights)
            e_i, e_s = oracle_search(c, "hybrid", q_bits[q], sp_terms, sp_weights, None, 10)
            assert_result_equal(ids[0], scores[0], int(counts[0]), e_i, e_s, ctx=f"INTEGRATION.md snippet, query {q}")
    shard.close()


def test_pipelined_searches_back_to_back(gpu, monkeypatch):
    """Pipelined mode (b200rag_set_pipeline through ShardedSearcher): 90 staged searches enqueued back to back with NO
    host synchronisation -- single queries (pipelined form: scan on the shard's stream, tails + fuse on the result
    stream, double-buffered candidate lists) interleaved with batches of 4 (tcgen05 path: classic form inside the
    pipeline, drain + hand-over) and sparse-only / dense-only searches.  Every step's fused output is copied aside on the
    result stream and compared with the oracle afterwards: a missing event or a reused buffer shows up as a stale or
    torn result."""
    import torch
    from b200rag import Shard, normalize_bf16
    from b200rag.dist import ShardedSearcher
    monkeypatch.setenv("B200RAG_PIPELINE_TAIL", "1")
    c = Corpus(50_000, dim=1024, vocab=60_013)
    dev = torch.device("cuda", gpu)
    sh = Shard(dim=1024, vocab=c.vocab, device=gpu, docs_per_block=2048)
    sh.add(c.bits, c.indptr, c.terms, c.w)
    ss = ShardedSearcher(sh, dev)
    assert ss.pipeline and ss.result_stream() is not None
    nst, k = 90, 10
    plan = []                                            # (mode, first query, batch)
    qn = 0
    for i in range(nst):
        mode = ("hybrid", "hybrid", "dense", "hybrid", "sparse")[i % 5]
        B = 4 if i % 7 == 3 else 1
        plan.append((mode, qn, B))
        qn += B
    qf, ip, tt, ww = c.queries(qn, qid_start=300)
    qb = normalize_bf16(qf)
    for i, (mode, q0, B) in enumerate(plan):
        ss.stage(mode, k, qb[q0:q0 + B], ip[q0:q0 + B + 1] - ip[q0], tt[ip[q0]:ip[q0 + B]], ww[ip[q0]:ip[q0 + B]], slot=i)
    keep = []
    for i in range(nst):
        ss.use_slot(i)
        b = ss.run_staged()
        with torch.cuda.stream(ss.result_stream()):
            keep.append(b["out"].clone())
    torch.cuda.synchronize(dev)
    for i, (mode, q0, B) in enumerate(plan):
        h = keep[i].cpu().numpy()
        ids = h[:B * k].reshape(B, k)
        sc = h[B * k:2 * B * k].view(np.float64).reshape(B, k)
        cnt = h[2 * B * k:].view(np.int32)
        assert cnt[B] == 0 and cnt[B + 1] == 0
        for b_ in range(B):
            q = q0 + b_
            e_i, e_s = oracle_search(c, mode, qb[q], tt[ip[q]:ip[q + 1]], ww[ip[q]:ip[q + 1]], None, k)
            assert_result_equal(ids[b_], sc[b_], int(cnt[b_]), e_i, e_s, ctx=f"pipelined step {i} ({mode}, B={B}) q{b_}")
    # the synchronous host-buffer call still works on a pipelined shard (classic form), and so does a retry-free search()
    r = ss.search("hybrid", k, qb[:1], ip[:2], tt[:ip[1]], ww[:ip[1]])
    e_i, e_s = oracle_search(c, "hybrid", qb[0], tt[:ip[1]], ww[:ip[1]], None, k)
    assert_result_equal(r[0][0], r[1][0], int(r[2][0]), e_i, e_s, ctx="search() on a pipelined searcher")
    r = sh.search("hybrid", k, qb[:1], ip[:2], tt[:ip[1]], ww[:ip[1]])
    assert_result_equal(r[0][0], r[1][0], int(r[2][0]), e_i, e_s, ctx="b200rag_search on a pipelined shard")
    sh.close()
