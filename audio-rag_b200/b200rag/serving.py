"""Serving-side plumbing for the retriever: take the blocking search off the event loop and turn
concurrent requests into batches.

The reference's FastAPI handler calls the blocking ``pipeline.query(...)`` straight from ``async def search_audio``
(/root/reference/src/audio_rag/api/v1/query.py:90-115), so one slow search stalls every other request of the worker.
``BatchingSearcher`` is what that handler (or ``QueryPipeline``) awaits instead:

    searcher = BatchingSearcher(retriever, max_batch=64, max_wait_ms=0.5)
    results = await searcher.search(query_embedding, top_k=10, collection_name="tenant_a", search_type="hybrid")

* every engine call runs on ONE worker thread (the engine handle is safe for serialised use from one thread at a time,
  like the reference's retriever: SURVEY 8b "Threading"), never on the event loop;
* requests that arrive while a call is running are collected and go out TOGETHER as one ``search_batch`` -- one pass over
  the corpus for up to ``max_batch`` queries (per-query collection names are supported by ``search_batch``; requests
  are grouped by (top_k, filter_metadata, search_type), the arguments ``search_batch`` shares across a batch);
* ``max_wait_ms`` bounds how long the first request of a batch waits for company when the engine is idle.

No reference counterpart (additive, like ``search_batch``).  Host-only code: no GPU is needed to test it.
"""
from __future__ import annotations

import asyncio
import threading
from concurrent.futures import ThreadPoolExecutor


def _key(top_k, filter_metadata, search_type):
    flt = tuple(sorted((str(k), repr(v)) for k, v in (filter_metadata or {}).items()))
    return (top_k, flt, search_type)


class BatchingSearcher:
    def __init__(self, retriever, max_batch: int = 64, max_wait_ms: float = 0.5):
        self.retriever = retriever
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="b200rag-search")
        self._pending: list = []          # (key, embedding, collection, kwargs, loop, future)
        self._lock = threading.Lock()
        self._draining = False
        self.batches: list[int] = []      # sizes of the engine calls made (observability / tests)

    async def search(self, query_embedding, top_k=None, collection_name=None, filter_metadata=None, search_type=None):
        """Awaitable twin of ``retriever.search`` (same arguments, same result, same ``RetrievalError``)."""
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        item = (_key(top_k, filter_metadata, search_type), query_embedding, collection_name,
                {"top_k": top_k, "filter_metadata": filter_metadata, "search_type": search_type}, loop, fut)
        with self._lock:
            self._pending.append(item)
            start = not self._draining
            if start:
                self._draining = True
        if start:
            if self.max_wait > 0:
                await asyncio.sleep(self.max_wait)      # let requests of the same tick join the first batch
            self._pool.submit(self._drain)
        return await fut

    async def search_batch(self, query_embeddings, **kw):
        """A caller-made batch goes through the same worker thread (serialised with everything else)."""
        loop = asyncio.get_running_loop()
        return await loop.run_in_executor(self._pool, lambda: self.retriever.search_batch(query_embeddings, **kw))

    def _drain(self):
        """Worker thread: keep taking what has queued up, one homogeneous batch at a time, until nothing is pending."""
        while True:
            with self._lock:
                if not self._pending:
                    self._draining = False
                    return
                key = self._pending[0][0]
                take = [it for it in self._pending if it[0] == key][:self.max_batch]
                ids = {id(it) for it in take}
                self._pending = [it for it in self._pending if id(it) not in ids]
            self.batches.append(len(take))
            try:
                if len(take) == 1:
                    _, emb, coll, kw, _, _ = take[0]
                    results = [self.retriever.search(emb, collection_name=coll, **kw)]
                else:
                    kw = take[0][3]
                    results = self.retriever.search_batch([it[1] for it in take], collection_name=[it[2] for it in take], **kw)
                for it, res in zip(take, results):
                    it[4].call_soon_threadsafe(_resolve, it[5], res, None)
            except Exception as e:           # RetrievalError for the whole batch: every waiter sees it
                for it in take:
                    it[4].call_soon_threadsafe(_resolve, it[5], None, e)

    def close(self):
        self._pool.shutdown(wait=True)


def _resolve(fut, result, error):
    if fut.done():
        return
    if error is not None:
        fut.set_exception(error)
    else:
        fut.set_result(result)
