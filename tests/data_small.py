"""Small deterministic chunk/embedding sets shared by the adapter tests and the golden-fixture generator."""
from __future__ import annotations

import numpy as np

from b200rag import synth

DIM = 256
VOCAB = 250_002


def make_chunks(n, seed, collection_tag, AudioChunk, EmbeddingResult, SparseVector, sparse=True, row_start=0, dim=DIM):
    dense = synth.dense_rows_f32(seed, row_start, n, dim)
    thr = synth.zipf_thresholds(VOCAB)
    ip, tt, ww = synth.sparse_docs_csr(seed, row_start, n, 10_000, VOCAB, 64, thr, synth.bm25_tables(10_000, VOCAB, 64))
    chunks, embs = [], []
    for i in range(n):
        meta = {"source": f"{collection_tag}-{i % 3}.wav", "lang": "en" if i % 2 == 0 else "de", "idx": i,
                "tags": ["a", "b"] if i % 4 == 0 else ["c"]}
        chunks.append(AudioChunk(text=f"{collection_tag} chunk {i}", start=float(i), end=float(i) + 0.5,
                                 speaker=f"SPEAKER_{i % 2:02d}" if i % 5 else None, metadata=meta))
        sv = None
        if sparse:
            # shuffled index order on purpose: the plugin must sort
            sl = slice(ip[i], ip[i + 1])
            perm = np.random.default_rng(seed * 1000 + i).permutation(ip[i + 1] - ip[i])
            sv = SparseVector(indices=[int(x) for x in tt[sl][perm]], values=[float(x) for x in ww[sl][perm]])
        embs.append(EmbeddingResult(dense=[float(x) for x in dense[i] * (1.0 + i % 7)], sparse=sv))
    return chunks, embs


def make_queries(nq, seed, n_rows, corpus_seed, EmbeddingResult, SparseVector, sparse=True, dim=DIM):
    qf = synth.dense_queries_f32(seed, 0, nq, n_rows, dim, corpus_seed=corpus_seed)
    qi, qt, qw = synth.sparse_queries(seed, 0, nq, 10, VOCAB)
    out = []
    for i in range(nq):
        sv = SparseVector(indices=[int(x) for x in qt[qi[i]:qi[i + 1]][::-1]],
                          values=[float(x) for x in qw[qi[i]:qi[i + 1]][::-1]]) if sparse else None
        out.append(EmbeddingResult(dense=[float(x) for x in qf[i]], sparse=sv))
    return out


def result_rows(results):
    """(text, score) view of a result list - texts encode the row, so this pins ids without uuids."""
    return [(r.chunk.text, r.score, r.source, r.chunk.start, r.chunk.end, r.chunk.speaker, r.chunk.metadata)
            for r in results]
