"""BatchingSearcher (b200rag/serving.py): searches leave the event loop, concurrent requests are
coalesced into search_batch calls, results and errors reach the right waiter.  Host-only (oracle double)."""
import asyncio
import threading
import time

import pytest

from data_small import DIM, make_chunks, make_queries, result_rows
from oracle_shard import OracleShard


def _types():
    from b200rag.compat import AudioChunk, EmbeddingResult, SparseVector
    return AudioChunk, EmbeddingResult, SparseVector


def _retriever():
    from b200rag.compat import RetrievalConfig
    from b200rag.retriever import B200Retriever
    try:
        conf = RetrievalConfig(qdrant_in_memory=True, top_k=5)
    except TypeError:
        conf = RetrievalConfig(top_k=5)
    r = B200Retriever(conf, embedding_dim=DIM)
    r._shard = OracleShard(dim=DIM)
    return r


def test_concurrent_requests_are_batched_off_the_event_loop():
    from b200rag.serving import BatchingSearcher
    A, E, S = _types()
    r = _retriever()
    for name, seed in (("a", 1), ("b", 2)):
        ch, em = make_chunks(120, seed, name.upper(), A, E, S)
        r.add(ch, em, name)
    qs = make_queries(12, 9, 120, 1, E, S)
    names = ["a", "b"] * 6
    expect = [result_rows(r.search(q, top_k=4, collection_name=n, search_type="hybrid")) for q, n in zip(qs, names)]
    seen_threads = set()
    orig_batch, orig_one = r.search_batch, r.search

    def spy_batch(*a, **kw):
        seen_threads.add(threading.get_ident())
        time.sleep(0.02)                       # a "slow" engine call: requests pile up behind it
        return orig_batch(*a, **kw)

    def spy_one(*a, **kw):
        seen_threads.add(threading.get_ident())
        time.sleep(0.02)
        return orig_one(*a, **kw)

    r.search_batch, r.search = spy_batch, spy_one
    bs = BatchingSearcher(r, max_batch=8, max_wait_ms=2.0)
    ticks = []

    async def heartbeat():
        for _ in range(20):
            ticks.append(time.perf_counter())
            await asyncio.sleep(0.002)

    async def main():
        hb = asyncio.create_task(heartbeat())
        got = await asyncio.gather(*[bs.search(q, top_k=4, collection_name=n, search_type="hybrid")
                                     for q, n in zip(qs, names)])
        # a different argument set forms its own batch; a failing request raises for its waiter only
        other = await bs.search(qs[0], top_k=2, collection_name="a", search_type="dense")
        await hb
        return got, other

    got, other = asyncio.run(main())
    assert [result_rows(x) for x in got] == expect
    assert len(other) == 2
    assert threading.get_ident() not in seen_threads and len(seen_threads) == 1, "engine calls run on ONE worker thread"
    assert max(bs.batches) > 1 and sum(bs.batches) == 13, bs.batches
    assert max(b for b in bs.batches) <= 8
    gaps = [b - a for a, b in zip(ticks, ticks[1:])]
    assert max(gaps) < 0.015, "the event loop kept ticking while searches ran"
    bs.close()


def test_errors_reach_the_waiter():
    from b200rag.compat import RetrievalError
    from b200rag.serving import BatchingSearcher
    A, E, S = _types()
    r = _retriever()
    ch, em = make_chunks(30, 3, "A", A, E, S)
    r.add(ch, em, "a")
    q = make_queries(1, 9, 30, 3, E, S)[0]
    bad = E(dense=[0.0] * (DIM // 2), sparse=None)
    bs = BatchingSearcher(r, max_wait_ms=0)

    async def main():
        with pytest.raises(RetrievalError):
            await bs.search(bad, collection_name="a")
        return await bs.search(q, collection_name="a")

    assert len(asyncio.run(main())) == 5
    bs.close()
