#!/usr/bin/env python
"""End-to-end through the PLUGIN with several shards in ONE process (B200Retriever(devices=[...]) -> b200rag_group_search):
wall clock of `B200Retriever.search(EmbeddingResult)` -> list[RetrievalResult], the call `AudioRAG.query()` makes.

    python tools/plugin_group_bench.py --devices 0,1,2,3,4,5,6,7 [--rows 10000000] [--steps 200]
    python tools/plugin_group_bench.py --devices 0,0                # two shards on one GPU (functional check)

The corpus is the bench's synthetic one, row-sharded over the devices (contiguous ranges, built on each GPU with the device
generators) and adopted with `attach_prebuilt`; results of the first queries are checked against a 1-shard search when
--check is given and the rows fit one device."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-rag_b200"), os.path.join(ROOT, "tools")]
from b200rag import synth  # noqa: E402
from b200rag.compat import EmbeddingResult, RetrievalConfig, SparseVector  # noqa: E402
from b200rag.dist import shard_bounds  # noqa: E402
from b200rag.retriever import B200Retriever  # noqa: E402
from probe import build_shard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0")
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--top-k", type=int, default=10)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    devs = [int(x) for x in a.devices.split(",")]
    n = len(devs)
    shards = []
    for i, d in enumerate(devs):
        lo, hi = shard_bounds(a.rows, n, i, align=8192)
        dev = torch.device("cuda", d)
        torch.cuda.set_device(dev)
        sh = build_shard(hi - lo, 1024, True, dev, n_total=a.rows, row0=lo)
        shards.append(sh)
    for d in set(devs):
        torch.cuda.synchronize(torch.device("cuda", d))
    try:
        conf = RetrievalConfig(qdrant_in_memory=True, top_k=a.top_k, search_type="hybrid")
    except TypeError:
        conf = RetrievalConfig(top_k=a.top_k, search_type="hybrid")
    r = B200Retriever(conf, embedding_dim=1024)
    r.attach_prebuilt(shards, "bench")
    nq = (a.warmup + a.steps) * a.batch
    qf = synth.dense_queries_f32(2000, 0, nq, a.rows, 1024, corpus_seed=1234)
    qi, qt, qw = synth.sparse_queries(2000, 0, nq)
    embs = [EmbeddingResult(dense=[float(x) for x in qf[i]],
                            sparse=SparseVector(indices=[int(t) for t in qt[qi[i]:qi[i + 1]]],
                                                values=[float(w) for w in qw[qi[i]:qi[i + 1]]])) for i in range(nq)]
    ts = []
    first = None
    for s in range(a.warmup + a.steps):
        batch = embs[s * a.batch:(s + 1) * a.batch]
        t0 = time.perf_counter()
        res = r.search(batch[0], collection_name="bench") if a.batch == 1 else r.search_batch(batch, collection_name="bench")
        if s >= a.warmup:
            ts.append(time.perf_counter() - t0)
        if s == 0:
            first = res
    out = {"what": "B200Retriever.search through b200rag_group_search" if n > 1 else "B200Retriever.search, one shard",
           "devices": devs, "rows": a.rows, "batch": a.batch, "top_k": a.top_k, "steps": a.steps,
           "queries_per_s": a.batch * len(ts) / float(np.sum(ts)), "p50_ms": float(np.median(ts) * 1e3),
           "p95_ms": float(np.percentile(ts, 95) * 1e3)}
    if a.check and n > 1:
        dev = torch.device("cuda", devs[0])
        torch.cuda.set_device(dev)
        one = build_shard(a.rows, 1024, True, dev)
        r1 = B200Retriever(conf, embedding_dim=1024)
        r1.attach_prebuilt([one], "bench")
        ref = r1.search(embs[0], collection_name="bench") if a.batch == 1 else r1.search_batch(embs[:a.batch], collection_name="bench")
        flat = lambda x: [(h.chunk.text, h.score) for h in (x if a.batch == 1 else [y for row in x for y in row])]   # noqa: E731
        out["equals_single_shard"] = flat(first) == flat(ref)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
