"""N > 1 plumbing on CPU: world_size-2 `gloo`, each rank holding a CPU double of its shard (oracle-backed), through
the same `ShardedSearcher` the GPU path uses (broadcast of the query, all-gather of candidate buffers with their
trailers, fused result identical on every rank and identical to a single-shard oracle)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class CpuShardDouble:
    """Duck-types the Shard calls ShardedSearcher makes; numpy restatement of legs/fuse (test infrastructure)."""

    def __init__(self, corpus, lo, hi, dim):
        from oracle import oracle
        self.o = oracle
        self.dim, self.row_base = dim, lo
        self.idx = oracle.OracleIndex(dim)
        ip = corpus.indptr[lo:hi + 1] - corpus.indptr[lo]
        self.idx.add_bits(corpus.bits[lo:hi], ip, corpus.terms[corpus.indptr[lo]:corpus.indptr[hi]],
                          corpus.w[corpus.indptr[lo]:corpus.indptr[hi]])
        self.slack_calls = []

    def make_query(self, mode, top_k, q_bits=None, sp_indptr=None, sp_terms=None, sp_weights=None, mask_ids=None,
                   score_threshold=None, rrf_k=0):
        from b200rag._ffi import Shard
        self._pending = dict(mode=mode, top_k=top_k, q_bits=q_bits, ip=sp_indptr, tt=sp_terms, ww=sp_weights)
        return Shard.make_query(self, mode, top_k, q_bits, sp_indptr, sp_terms, sp_weights, mask_ids,
                                score_threshold, rrf_k)

    def stage(self, q, keep=None, slot=0):
        self._slots = getattr(self, "_slots", {})
        self._slots[slot] = self._pending
        self._staged = self._pending

    def use_slot(self, slot):
        self._staged = self._slots[slot]

    def set_slack(self, s):
        self.slack_calls.append(s)

    def legs(self, mine, amb):
        from b200rag._ffi import CAND_DTYPE
        s = self._staged
        mode, k = s["mode"], s["top_k"]
        B = s["q_bits"].shape[0]
        L = 2 * k if mode == "hybrid" else k
        nlegs = 2 if mode == "hybrid" else 1
        c = mine.numpy().view(CAND_DTYPE).reshape(-1)[:nlegs * B * L].reshape(nlegs, B, L)
        c[...] = np.zeros((), CAND_DTYPE)
        elig = np.ones(self.idx.n, bool)
        for b in range(B):
            legs = []
            if mode != "sparse":
                legs.append(self.idx.dense_leg(s["q_bits"][b], elig, L))
            if mode != "dense":
                sl = slice(s["ip"][b], s["ip"][b + 1])
                legs.append(self.idx.sparse_leg(s["tt"][sl], s["ww"][sl], elig, L))
            for li, (ids, sc) in enumerate(legs):
                c["id"][li, b, :len(ids)] = ids + self.row_base
                c["score"][li, b, :len(ids)] = sc
                c["valid"][li, b, :len(ids)] = 1

    def fuse(self, src, world, out_ids, out_scores, out_counts, has_trailer=False):
        from b200rag._ffi import CAND_DTYPE
        s = self._staged
        mode, k = s["mode"], s["top_k"]
        B = s["q_bits"].shape[0]
        L = 2 * k if mode == "hybrid" else k
        nlegs = 2 if mode == "hybrid" else 1
        g = src.numpy().view(CAND_DTYPE).reshape(world, -1)
        body = g[:, :nlegs * B * L].reshape(world, nlegs, B, L)
        oi, os_ = out_ids.numpy().reshape(B, k), out_scores.numpy().view(np.float64).reshape(B, k)
        oc = out_counts.numpy().view(np.int32)
        oc[B] = int(g["id"][:, -1].sum())
        for b in range(B):
            merged = []
            for li in range(nlegs):
                e = body[:, li, b, :].reshape(-1)
                e = e[e["valid"] == 1]
                order = np.lexsort((e["id"], -e["score"].astype(np.float64)))[:L]
                merged.append(e[order])
            if nlegs == 1:
                ids, sc = merged[0]["id"][:k], merged[0]["score"][:k].astype(np.float64)
            else:
                ids, sc = self.o.rrf_fuse([merged[0]["id"], merged[1]["id"]], k)
            oc[b] = len(ids)
            oi[b, :len(ids)] = ids
            os_[b, :len(ids)] = sc


def _worker(rank, world, port, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "audio-rag_b200"), os.path.join(ROOT, "tests")):
            sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from b200rag.dist import ShardedSearcher, shard_bounds
        from helpers import Corpus, oracle_search
        from oracle import oracle
        n, dim = 3000, 256
        c = Corpus(n, dim=dim, vocab=20_011)
        lo, hi = shard_bounds(n, world, rank, align=16)
        assert (lo % 16 == 0) and sum(1 for _ in range(1)) == 1
        ss = ShardedSearcher(CpuShardDouble(c, lo, hi, dim), torch.device("cpu"))
        qf, ip, tt, ww = c.queries(4)
        qb = oracle.normalize_bf16(qf)
        # serving shape: only rank 0 holds the request, everyone gets it by broadcast
        arrays = ss.broadcast_query({"qb": qb, "ip": ip, "tt": tt, "ww": ww} if rank == 0 else None, src=0)
        assert np.array_equal(arrays["qb"], qb)
        for mode, k in (("dense", 10), ("sparse", 7), ("hybrid", 10), ("hybrid", 3)):
            ids, sc, cnt = ss.search(mode, k, arrays["qb"], arrays["ip"], arrays["tt"], arrays["ww"])
            for b in range(4):
                sl = slice(ip[b], ip[b + 1])
                ei, es = oracle_search(c, mode, qb[b], tt[sl], ww[sl], None, k)
                assert cnt[b] == len(ei) and np.array_equal(ids[b, :cnt[b]], ei), (rank, mode, b)
                assert np.array_equal(sc[b, :cnt[b]], es), (rank, mode, b)
            # identical on every rank
            gathered = [None] * world
            dist.all_gather_object(gathered, ids.tolist())
            assert all(g == gathered[0] for g in gathered)
        # a queue of staged batches replayed out of order (the bench's pre-staged slots): stage 3 single queries,
        # run them 2, 0, 1 and compare with the oracle
        k = 5
        for i in range(3):
            ss.stage("hybrid", k, qb[i:i + 1], ip[i:i + 2] - ip[i], tt[ip[i]:ip[i + 1]], ww[ip[i]:ip[i + 1]], slot=i)
        for i in (2, 0, 1):
            ss.use_slot(i)
            ids, sc, cnt, amb = ss.fetch(ss.run_staged())
            ei, es = oracle_search(c, "hybrid", qb[i], tt[ip[i]:ip[i + 1]], ww[ip[i]:ip[i + 1]], None, k)
            assert amb == 0 and cnt[0] == len(ei) and np.array_equal(ids[0, :cnt[0]], ei) and np.array_equal(sc[0, :cnt[0]], es)
        assert not ss.p2p            # CPU doubles never take the peer-memory path
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))


def test_shard_bounds_cover_everything():
    from b200rag.dist import shard_bounds
    for n, w, a in ((10_000_000, 8, 8192), (1_000_003, 4, 1), (100, 8, 16), (5, 8, 1)):
        spans = [shard_bounds(n, w, r, a) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert all(lo % a == 0 or lo == n for lo, _ in spans)


@pytest.mark.timeout(300)
def test_two_rank_gloo_search_matches_single_shard_oracle():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=280) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
