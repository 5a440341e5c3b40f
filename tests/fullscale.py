"""Full-corpus oracle at BASELINE sizes (TEST INFRASTRUCTURE; also used by bench.py's `oracle_check`, outside the
timed region).

The shard's STORED rows (bf16 bits, forward sparse index: `read_dense` / `read_sparse`) are streamed back to the host
in chunks and every row is scored by the C oracle (`oracle/oracle_c.c`: canonical fp64 sums, rules R2/R3/R7); the
per-leg top-L under R4-R7 is kept with an exact running merge (score desc, smaller id first).  Nothing here looks at
what the engine returned: the result is the oracle's answer over ALL rows, so a row the scan kernels missed shows up as
a mismatch (the former spot checks would have passed it).  10M rows x 8 queries take ~40 s on the GPU box's host cores.
"""
from __future__ import annotations

import numpy as np

from oracle import fast, oracle


def topk_exact(scores, ids, eligible, limit):
    """Exact top-`limit` of (scores, ids) restricted to `eligible` under R5 (score desc, id asc) without sorting the
    whole chunk: everything at or above the limit-th largest score is a candidate, then one small lexsort."""
    sel = np.flatnonzero(eligible)
    if len(sel) == 0 or limit <= 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float32)
    s = scores[sel]
    if len(sel) > limit:
        kth = np.partition(s, len(s) - limit)[len(s) - limit]
        keep = s >= kth
        sel, s = sel[keep], s[keep]
    i = ids[sel]
    order = np.lexsort((i, -s.astype(np.float64)))[:limit]
    return i[order].astype(np.int64), s[order].astype(np.float32)


def merge_topk(a, b, limit):
    i = np.concatenate([a[0], b[0]])
    s = np.concatenate([a[1], b[1]])
    order = np.lexsort((i, -s.astype(np.float64)))[:limit]
    return i[order], s[order]


class LegJob:
    """One query whose exact per-leg top-L over the whole shard is wanted.  `eligible(lo, hi)` -> bool[hi - lo] for
    LOCAL rows [lo, hi), or None = every row."""

    def __init__(self, q_bits=None, q_idx=None, q_val=None, L=20, eligible=None):
        self.q_bits, self.q_idx, self.q_val, self.L, self.eligible = q_bits, q_idx, q_val, L, eligible
        e = (np.zeros(0, np.int64), np.zeros(0, np.float32))
        self.dense, self.sparse = e, e

    def hybrid(self, top_k, rrf_k=2):
        return oracle.rrf_fuse([self.dense[0][:2 * top_k], self.sparse[0][:2 * top_k]], top_k, rrf_k)


def stream_oracle_legs(shard, jobs, chunk=1 << 19, check_rows=None):
    """Fill job.dense / job.sparse = (global ids, fp32 scores) of the exact top-L of each leg over ALL rows of `shard`.
    `check_rows(lo, bits, indptr, terms, weights)` (optional) sees every chunk, e.g. to compare it with the generators."""
    n = shard.count
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        ids = shard.read_row_ids(lo, m)
        want_dense = any(j.q_bits is not None for j in jobs)
        want_sparse = any(j.q_idx is not None for j in jobs)
        bits = shard.read_dense(lo, m) if want_dense else None
        ip = tt = ww = None
        if want_sparse:
            ip, tt, ww = shard.read_sparse(lo, m)
        if check_rows is not None:
            check_rows(lo, bits, ip, tt, ww)
        for j in jobs:
            elig = np.ones(m, dtype=bool) if j.eligible is None else np.asarray(j.eligible(lo, lo + m), dtype=bool)
            if j.q_bits is not None:
                s = fast.dense_scores(bits, j.q_bits)
                j.dense = merge_topk(j.dense, topk_exact(s, ids, elig, j.L), j.L)
            if j.q_idx is not None:
                s, touched = fast.sparse_scores(ip, tt, ww, j.q_idx, j.q_val)
                j.sparse = merge_topk(j.sparse, topk_exact(s, ids, elig & touched, j.L), j.L)
    return jobs
