"""TEST DOUBLE with the method surface of ``b200rag._ffi.Shard`` that ``B200Retriever`` uses, backed by the CPU
oracle.  Tests inject it (``retriever._shard = OracleShard(...)``) to exercise the adapter's host logic without
a GPU and, on a GPU, to differential-test the real shard against the oracle through the very same adapter code.
The product never imports this module."""
from __future__ import annotations

import numpy as np

from oracle import oracle


class OracleShard:
    def __init__(self, dim=1024, vocab=250_002, device=0, row_base=0, docs_per_block=0, **kw):
        self.dim, self.vocab, self.row_base = dim, vocab, row_base
        self.index = oracle.OracleIndex(dim)
        self.ids = np.zeros(0, dtype=np.int64)          # global id of every local row (strictly increasing)
        self.masks: dict[int, np.ndarray] = {}
        self.calls: list = []

    def add(self, dense_bits, sp_indptr=None, sp_terms=None, sp_weights=None, ids=None):
        n0 = self.index.n
        self.index.add_bits(dense_bits, sp_indptr, sp_terms, sp_weights)
        n = self.index.n - n0
        new = np.arange(self.row_base + n0, self.row_base + n0 + n, dtype=np.int64) if ids is None else \
            np.asarray(ids, dtype=np.int64)
        if len(new) != n or (len(new) > 1 and (np.diff(new) <= 0).any()) or \
                (len(self.ids) and n and new[0] <= self.ids[-1]):
            raise RuntimeError("add: row ids must be strictly increasing over the shard's lifetime")
        self.ids = np.concatenate([self.ids, new])

    def add_f32(self, dense_f32, sp_indptr=None, sp_terms=None, sp_weights=None, ids=None):
        self.add(oracle.normalize_bf16(dense_f32), sp_indptr, sp_terms, sp_weights, ids)

    def compact(self, keep_words, n_rows):
        if n_rows != self.index.n:
            raise RuntimeError("compact: the keep mask must cover exactly the shard's rows")
        keep = np.unpackbits(np.asarray(keep_words, dtype=np.uint32).view(np.uint8), bitorder="little")[:n_rows].astype(bool)
        ix = self.index
        lens = np.diff(ix.indptr)
        pos = np.repeat(keep, lens)
        ix.terms, ix.weights = ix.terms[pos], ix.weights[pos]
        ix.indptr = np.concatenate([[0], np.cumsum(lens[keep])]).astype(np.int64)
        ix.bits = ix.bits[keep]
        self.ids = self.ids[keep]
        self.masks.clear()

    def set_compression(self, on=True):
        # the 8-bit candidate scan changes where candidates come from, never the result: nothing to model, only to record
        if on and self.dim not in (512, 1024):
            raise RuntimeError("set_compression: the 8-bit scan needs dim 512 or 1024")
        self.calls.append(("set_compression", bool(on)))

    def clear(self):
        self.index = oracle.OracleIndex(self.dim)
        self.ids = np.zeros(0, dtype=np.int64)
        self.masks.clear()

    def close(self):
        pass

    # persistence twin of b200rag_save / b200rag_load (same refusal rules: empty shard, matching geometry)
    def save(self, path):
        ix = self.index
        np.savez(path + ".npz", dim=self.dim, vocab=self.vocab, bits=ix.bits, indptr=ix.indptr, terms=ix.terms,
                 weights=ix.weights, ids=self.ids)
        import os
        os.replace(path + ".npz", path)

    def load(self, path):
        if self.index.n:
            raise RuntimeError("load: the shard must be empty")
        try:
            z = np.load(path, allow_pickle=False)
            if int(z["dim"]) != self.dim or int(z["vocab"]) != self.vocab:
                raise RuntimeError("load: file was written for another dim/vocab")
        except (ValueError, OSError, KeyError) as e:
            raise RuntimeError(f"load: not a shard file ({e})")
        ix = self.index
        ix.bits, ix.indptr, ix.terms, ix.weights = z["bits"], z["indptr"], z["terms"], z["weights"]
        self.ids = z["ids"]

    @property
    def count(self):
        return self.index.n

    def mask_set(self, mask_id, words, n_rows):
        bits = np.unpackbits(np.asarray(words, dtype=np.uint32).view(np.uint8), bitorder="little")[:n_rows]
        self.masks[mask_id] = bits.astype(bool)

    def mask_drop(self, mask_id):
        self.masks.pop(mask_id, None)

    def _elig(self, mask_ids, b):
        elig = np.ones(self.index.n, dtype=bool)
        if mask_ids is not None and mask_ids[b] >= 0:
            m = self.masks[int(mask_ids[b])]
            elig = np.zeros(self.index.n, dtype=bool)
            elig[:len(m)] = m[:self.index.n]
        return elig

    def legs(self, mode, b, L, q_bits, sp_indptr, sp_terms, sp_weights, mask_ids, score_threshold):
        """Per-leg (global ids, fp32 scores) of query b over this shard's rows: what `b200rag_legs` emits."""
        elig = self._elig(mask_ids, b)
        out = []
        if mode != "sparse":
            i, s = self.index.dense_leg(q_bits[b], elig, L, score_threshold if mode == "dense" else None)
            out.append((self.ids[i], s))
        if mode != "dense":
            qi = sp_terms[sp_indptr[b]:sp_indptr[b + 1]]
            qv = sp_weights[sp_indptr[b]:sp_indptr[b + 1]]
            i, s = self.index.sparse_leg(qi, qv, elig, L) if self.index.n else (np.zeros(0, np.int64), np.zeros(0, np.float32))
            out.append((self.ids[i], s))
        return out

    def search(self, mode, top_k, q_bits=None, sp_indptr=None, sp_terms=None, sp_weights=None, mask_ids=None,
               score_threshold=None, rrf_k=0):
        return OracleGroup([self]).search(mode, top_k, q_bits, sp_indptr, sp_terms, sp_weights, mask_ids,
                                          score_threshold, rrf_k)


class OracleGroup:
    """TEST DOUBLE of ``b200rag._ffi.ShardGroup``: per-shard oracle legs, merged under R5, fused under R9/R10."""

    def __init__(self, shards):
        self.shards = list(shards)
        self.dim = self.shards[0].dim

    def close(self):
        pass

    def search(self, mode, top_k, q_bits=None, sp_indptr=None, sp_terms=None, sp_weights=None, mask_ids=None,
               score_threshold=None, rrf_k=0):
        q_bits = np.asarray(q_bits, dtype=np.uint16).reshape(-1, self.dim)
        B = q_bits.shape[0]
        ids = np.full((B, top_k), -1, dtype=np.int64)
        scores = np.zeros((B, top_k), dtype=np.float64)
        counts = np.zeros(B, dtype=np.int32)
        L = 2 * top_k if mode == "hybrid" else top_k
        for sh in self.shards:
            sh.calls.append((mode, top_k, B))
        for b in range(B):
            per = [sh.legs(mode, b, L, q_bits, sp_indptr, sp_terms, sp_weights, mask_ids, score_threshold)
                   for sh in self.shards]
            merged = []
            for leg in range(len(per[0])):
                i = np.concatenate([p[leg][0] for p in per])
                s = np.concatenate([p[leg][1] for p in per])
                order = np.lexsort((i, -s.astype(np.float64)))[:L]
                merged.append((i[order], s[order]))
            if mode == "hybrid":
                ri, rs = oracle.rrf_fuse([merged[0][0], merged[1][0]], top_k, rrf_k or oracle.RRF_K)
            else:
                ri, rs = merged[0][0], merged[0][1].astype(np.float64)
            counts[b] = len(ri)
            ids[b, :len(ri)] = ri
            scores[b, :len(ri)] = rs
        return ids, scores, counts
