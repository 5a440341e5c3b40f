"""Shared builders for the parity tests: synthetic corpora (host twins of the device generators),
oracle-side expected results."""
from __future__ import annotations

import numpy as np

from b200rag import synth
from oracle import fast, oracle

SEED = 1234
QSEED = 2000


class Corpus:
    def __init__(self, n, dim=1024, vocab=synth.VOCAB, doc_tokens=256, seed=SEED, n_total=None, row_start=0,
                 sparse=True, use_c=True):
        self.n, self.dim, self.vocab, self.seed, self.row_start = n, dim, vocab, seed, row_start
        n_total = n if n_total is None else n_total
        self.n_total = n_total
        gen_dense = fast.synth_dense_bf16 if use_c else synth.dense_rows_bf16
        self.bits = gen_dense(seed, row_start, n, dim)
        self.thr = synth.zipf_thresholds(vocab)
        self.idf, self.tff = synth.bm25_tables(n_total, vocab, doc_tokens)
        if sparse:
            if use_c:
                self.indptr, self.terms, self.w = fast.synth_sparse_csr(seed, row_start, n, self.thr, self.idf,
                                                                         self.tff, vocab, doc_tokens,
                                                                         synth.TERM_PERM_MUL % vocab)
            else:
                self.indptr, self.terms, self.w = synth.sparse_docs_csr(seed, row_start, n, n_total, vocab,
                                                                        doc_tokens, self.thr, (self.idf, self.tff))
        else:
            self.indptr = np.zeros(n + 1, dtype=np.int64)
            self.terms = np.zeros(0, np.uint32)
            self.w = np.zeros(0, np.float32)

    def queries(self, nq, qid_start=0, n_tokens=12, qseed=QSEED):
        qf = synth.dense_queries_f32(qseed, qid_start, nq, self.n_total, self.dim, corpus_seed=self.seed)
        ip, tt, ww = synth.sparse_queries(qseed, qid_start, nq, n_tokens, self.vocab, self.thr)
        return qf, ip, tt, ww


def oracle_search(c: Corpus, mode, q_bits, q_idx, q_val, eligible, top_k, score_threshold=None, rrf_k=2,
                  row_base=0):
    """Canonical oracle on the corpus arrays (C accelerated scoring, numpy selection/fusion)."""
    n = c.n
    elig = np.ones(n, dtype=bool) if eligible is None else np.asarray(eligible, dtype=bool)

    def dense_leg(limit, thr=None):
        s = fast.dense_scores(c.bits, q_bits) if n else np.zeros(0, np.float32)
        return oracle.leg_topk(s, elig, limit, thr)

    def sparse_leg(limit):
        if n == 0:
            return np.zeros(0, np.int64), np.zeros(0, np.float32)
        s, touched = fast.sparse_scores(c.indptr, c.terms, c.w, q_idx, q_val)
        return oracle.leg_topk(s, elig & touched, limit)

    if mode == "dense":
        i, s = dense_leg(top_k, score_threshold)
        return i + row_base, s.astype(np.float64)
    if mode == "sparse":
        i, s = sparse_leg(top_k)
        return i + row_base, s.astype(np.float64)
    di, _ = dense_leg(2 * top_k)
    si, _ = sparse_leg(2 * top_k)
    return oracle.rrf_fuse([di + row_base, si + row_base], top_k, rrf_k)


def assert_result_equal(got_ids, got_scores, got_count, exp_ids, exp_scores, ctx=""):
    assert got_count == len(exp_ids), f"{ctx}: count {got_count} != {len(exp_ids)}"
    gi = np.asarray(got_ids[:got_count])
    assert np.array_equal(gi, exp_ids), f"{ctx}: ids differ\n got {gi}\n exp {exp_ids}"
    gs = np.asarray(got_scores[:got_count], dtype=np.float64)
    assert np.array_equal(gs, np.asarray(exp_scores, dtype=np.float64)), \
        f"{ctx}: scores differ\n got {gs}\n exp {exp_scores}"
