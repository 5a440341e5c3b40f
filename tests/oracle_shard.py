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
        self.masks: dict[int, np.ndarray] = {}
        self.calls: list = []

    def add(self, dense_bits, sp_indptr=None, sp_terms=None, sp_weights=None):
        self.index.add_bits(dense_bits, sp_indptr, sp_terms, sp_weights)

    def clear(self):
        self.index = oracle.OracleIndex(self.dim)
        self.masks.clear()

    def close(self):
        pass

    # persistence twin of b200rag_save / b200rag_load (same refusal rules: empty shard, matching geometry)
    def save(self, path):
        ix = self.index
        np.savez(path + ".npz", dim=self.dim, vocab=self.vocab, bits=ix.bits, indptr=ix.indptr, terms=ix.terms,
                 weights=ix.weights)
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

    @property
    def count(self):
        return self.index.n

    def mask_set(self, mask_id, words, n_rows):
        bits = np.unpackbits(np.asarray(words, dtype=np.uint32).view(np.uint8), bitorder="little")[:n_rows]
        self.masks[mask_id] = bits.astype(bool)

    def mask_drop(self, mask_id):
        self.masks.pop(mask_id, None)

    def search(self, mode, top_k, q_bits=None, sp_indptr=None, sp_terms=None, sp_weights=None, mask_ids=None,
               score_threshold=None, rrf_k=0):
        q_bits = np.asarray(q_bits, dtype=np.uint16).reshape(-1, self.dim)
        B = q_bits.shape[0]
        ids = np.full((B, top_k), -1, dtype=np.int64)
        scores = np.zeros((B, top_k), dtype=np.float64)
        counts = np.zeros(B, dtype=np.int32)
        self.calls.append((mode, top_k, B))
        for b in range(B):
            elig = np.ones(self.index.n, dtype=bool)
            if mask_ids is not None and mask_ids[b] >= 0:
                m = self.masks[int(mask_ids[b])]
                elig = np.zeros(self.index.n, dtype=bool)
                elig[:len(m)] = m[:self.index.n]
            qi = qv = None
            if mode != "dense":
                qi = sp_terms[sp_indptr[b]:sp_indptr[b + 1]]
                qv = sp_weights[sp_indptr[b]:sp_indptr[b + 1]]
            i, s = self.index.search(mode, q_bits[b], qi, qv, elig, top_k, score_threshold, rrf_k or oracle.RRF_K)
            counts[b] = len(i)
            ids[b, :len(i)] = i + self.row_base
            scores[b, :len(i)] = s
        return ids, scores, counts
