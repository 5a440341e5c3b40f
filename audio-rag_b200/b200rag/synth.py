"""Deterministic synthetic corpus / query generators (host side, numpy).

Every generator here has a bit-identical CUDA twin in ``csrc/synth.cu`` (same
integer hash, same integer thresholds, same IEEE fp64 division), so a shard of
any row range can be produced on the GPU at full scale and on the CPU at test
scale with the *same bits*.  Spec: SURVEY.md §8d (shapes: 1024-d BGE-M3-like
unit vectors, 256-token chunks, Zipf(s=1) terms over the XLM-R index range
V = 250 002, BM25 impact weights k1 = 1.2, b = 0.75).

Nothing in here is floating-point order dependent:
  * the raw dense element is an integer (sum of four 16-bit hash fields,
    centred), the squared norm is an exact int64, the normalised value is
    ``fp32(x / sqrt(fp64(ss)))`` (one IEEE sqrt, one IEEE division, one cast)
    and the stored value is its bf16 round-to-nearest-even;
  * Zipf draws compare a 53-bit integer against an integer threshold table;
  * BM25 impacts are one fp32 multiply of two host-computed fp32 tables.
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64
GOLD = U64(0x9E3779B97F4A7C15)
M1 = U64(0xBF58476D1CE4E5B9)
M2 = U64(0x94D049BB133111EB)
ROWMUL = U64(0xD6E8FEB86659FD93)

STREAM_DENSE = 1
STREAM_QNOISE = 2
STREAM_DOC = 3
STREAM_QSPARSE = 4
STREAM_COLL = 5
STREAM_QPICK = 6

VOCAB = 250_002          # XLM-R index range emitted by BGE-M3 (embeddings/bge.py:95-102)
DOC_TOKENS = 256         # chunking default max_tokens (config/schema.py:37)
TERM_PERM_MUL = 100_003  # rank -> term id scatter; coprime to VOCAB
BM25_K1 = 1.2
BM25_B = 0.75


def mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(z, dtype=U64) + GOLD).astype(U64)
        z = ((z ^ (z >> U64(30))) * M1).astype(U64)
        z = ((z ^ (z >> U64(27))) * M2).astype(U64)
        return (z ^ (z >> U64(31))).astype(U64)


def stream_key(seed: int, stream: int) -> np.uint64:
    with np.errstate(over="ignore"):
        return mix64(np.array([(int(seed) * 0x10000 + int(stream)) & 0xFFFFFFFFFFFFFFFF], dtype=U64))[0]


def row_keys(skey, rows):
    with np.errstate(over="ignore"):
        rows = np.asarray(rows, dtype=U64)
        return mix64(U64(skey) ^ (rows * ROWMUL).astype(U64))


def _elems(rkeys, n):
    """hash(rkey + j) for j in [0, n) -> uint64 [len(rkeys), n]."""
    with np.errstate(over="ignore"):
        j = np.arange(n, dtype=U64)[None, :]
        return mix64((np.asarray(rkeys, dtype=U64)[:, None] + j).astype(U64))


def _raw_int(rkeys, dim):
    """Centred sum of four 16-bit fields: an Irwin-Hall(4) bell, int64 in [-131070, 131070]."""
    h = _elems(rkeys, dim)
    m = U64(0xFFFF)
    s = (h & m) + ((h >> U64(16)) & m) + ((h >> U64(32)) & m) + ((h >> U64(48)) & m)
    return s.astype(np.int64) - 131070


def f32_to_bf16_bits(y: np.ndarray) -> np.ndarray:
    """IEEE round-to-nearest-even fp32 -> bf16 (finite inputs)."""
    u = np.ascontiguousarray(y, dtype=np.float32).view(np.uint32)
    r = u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    return (r >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def _normalise_int(raw: np.ndarray) -> np.ndarray:
    """raw int64 [n, dim] -> fp32 unit rows: fp32(x / sqrt(fp64(sum x^2)))."""
    ss = (raw * raw).sum(axis=1)                      # exact int64
    ss = np.where(ss == 0, 1, ss)
    nrm = np.sqrt(ss.astype(np.float64))
    return (raw.astype(np.float64) / nrm[:, None]).astype(np.float32)


def dense_rows_f32(seed: int, row_start: int, n: int, dim: int = 1024) -> np.ndarray:
    rk = row_keys(stream_key(seed, STREAM_DENSE), np.arange(row_start, row_start + n))
    return _normalise_int(_raw_int(rk, dim))


def dense_rows_bf16(seed: int, row_start: int, n: int, dim: int = 1024) -> np.ndarray:
    """Corpus rows as stored by the engine: bf16 bits, uint16 [n, dim]."""
    out = np.empty((n, dim), dtype=np.uint16)
    step = 8192
    for s in range(0, n, step):
        e = min(n, s + step)
        out[s:e] = f32_to_bf16_bits(dense_rows_f32(seed, row_start + s, e - s, dim))
    return out


def query_target_rows(seed: int, qids, n_rows: int) -> np.ndarray:
    k = row_keys(stream_key(seed, STREAM_QPICK), np.asarray(qids))
    return (k % U64(n_rows)).astype(np.int64)


def dense_queries_f32(seed: int, qid_start: int, nq: int, n_rows: int, dim: int = 1024,
                      corpus_seed: int | None = None) -> np.ndarray:
    """Planted queries: 2*row_j + noise (sigma = 0.5), every 10th query pure noise.

    Returned as fp32 unit vectors -- what an embedder hands to ``search``."""
    corpus_seed = seed if corpus_seed is None else corpus_seed
    qids = np.arange(qid_start, qid_start + nq)
    noise = _raw_int(row_keys(stream_key(seed, STREAM_QNOISE), qids), dim)
    tgt = query_target_rows(seed, qids, n_rows)
    base = _raw_int(row_keys(stream_key(corpus_seed, STREAM_DENSE), tgt), dim)
    planted = (qids % 10 != 9).astype(np.int64)[:, None]
    return _normalise_int(2 * base * planted + noise)


# ---------------------------------------------------------------- sparse side

def zipf_thresholds(n: int, s: float = 1.0) -> np.ndarray:
    """Integer CDF thresholds: draw u in [0, 2^53) -> rank = #thresholds <= u."""
    p = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    t = np.floor(cdf * float(1 << 53)).astype(np.uint64)
    t[-1] = U64(1 << 53)
    return t


def zipf_rank_probs(n: int, s: float = 1.0) -> np.ndarray:
    p = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    return p / p.sum()


def rank_to_term(rank, vocab: int = VOCAB):
    return ((np.asarray(rank, dtype=np.int64) * TERM_PERM_MUL) % vocab).astype(np.int64)


def bm25_tables(n_docs_total: int, vocab: int = VOCAB, doc_tokens: int = DOC_TOKENS,
                s: float = 1.0) -> tuple[np.ndarray, np.ndarray]:
    """(idf32[vocab] indexed by TERM id, tff32[doc_tokens+1] indexed by tf).

    df is the analytic expectation N*(1-(1-p)^T) so that any shard can be
    generated without a global pass; dl == avgdl == doc_tokens."""
    p = zipf_rank_probs(vocab, s)
    df = n_docs_total * (1.0 - np.power(1.0 - p, doc_tokens))
    idf_rank = np.log(1.0 + (n_docs_total - df + 0.5) / (df + 0.5))
    idf = np.empty(vocab, dtype=np.float32)
    idf[rank_to_term(np.arange(vocab), vocab)] = idf_rank.astype(np.float32)
    tf = np.arange(doc_tokens + 1, dtype=np.float64)
    tff = (tf * (BM25_K1 + 1.0) / (tf + BM25_K1 * (1.0 - BM25_B + BM25_B * 1.0))).astype(np.float32)
    return idf, tff


def _draw_terms(rkeys, n_tokens: int, thresholds: np.ndarray, vocab: int) -> np.ndarray:
    u = _elems(rkeys, n_tokens) >> U64(11)
    rank = np.searchsorted(thresholds, u, side="right")
    rank = np.minimum(rank, vocab - 1)
    return rank_to_term(rank, vocab)


def sparse_docs_csr(seed: int, row_start: int, n: int, n_docs_total: int, vocab: int = VOCAB,
                    doc_tokens: int = DOC_TOKENS, thresholds: np.ndarray | None = None,
                    tables: tuple[np.ndarray, np.ndarray] | None = None):
    """Doc-major CSR (indptr int64[n+1], terms uint32 ascending per doc, weights fp32)."""
    thresholds = zipf_thresholds(vocab) if thresholds is None else thresholds
    idf, tff = bm25_tables(n_docs_total, vocab, doc_tokens) if tables is None else tables
    indptr = np.zeros(n + 1, dtype=np.int64)
    terms_l, w_l = [], []
    step = 4096
    for s in range(0, n, step):
        e = min(n, s + step)
        rk = row_keys(stream_key(seed, STREAM_DOC), np.arange(row_start + s, row_start + e))
        t = np.sort(_draw_terms(rk, doc_tokens, thresholds, vocab), axis=1)
        first = np.ones_like(t, dtype=bool)
        first[:, 1:] = t[:, 1:] != t[:, :-1]
        # run lengths = tf
        flat_first = first.ravel()
        starts = np.flatnonzero(flat_first)
        ends = np.append(starts[1:], t.size)
        # a run never crosses a row because column 0 is always a run start
        tf = (ends - starts).astype(np.int64)
        tt = t.ravel()[starts]
        indptr[s + 1:e + 1] = first.sum(axis=1)
        terms_l.append(tt.astype(np.uint32))
        w_l.append((idf[tt] * tff[tf]).astype(np.float32))
    indptr = np.cumsum(indptr)
    terms = np.concatenate(terms_l) if terms_l else np.zeros(0, np.uint32)
    w = np.concatenate(w_l) if w_l else np.zeros(0, np.float32)
    return indptr, terms, w


def sparse_queries(seed: int, qid_start: int, nq: int, n_tokens: int = 12, vocab: int = VOCAB,
                   thresholds: np.ndarray | None = None):
    """Query sparse vectors: (indptr int64[nq+1], terms uint32 ascending, weights fp32 = query tf)."""
    thresholds = zipf_thresholds(vocab) if thresholds is None else thresholds
    rk = row_keys(stream_key(seed, STREAM_QSPARSE), np.arange(qid_start, qid_start + nq))
    t = np.sort(_draw_terms(rk, n_tokens, thresholds, vocab), axis=1)
    indptr = [0]
    terms, w = [], []
    for i in range(nq):
        u, c = np.unique(t[i], return_counts=True)
        terms.append(u.astype(np.uint32))
        w.append(c.astype(np.float32))
        indptr.append(indptr[-1] + len(u))
    return (np.asarray(indptr, dtype=np.int64),
            np.concatenate(terms) if terms else np.zeros(0, np.uint32),
            np.concatenate(w) if w else np.zeros(0, np.float32))


def row_collections(seed: int, row_start: int, n: int, n_collections: int,
                    thresholds: np.ndarray | None = None) -> np.ndarray:
    """Row -> collection id, Zipf(s=1) over collection ids (skewed tenants)."""
    thresholds = zipf_thresholds(n_collections) if thresholds is None else thresholds
    rk = row_keys(stream_key(seed, STREAM_COLL), np.arange(row_start, row_start + n))
    u = mix64(rk) >> U64(11)
    c = np.searchsorted(thresholds, u, side="right")
    return np.minimum(c, n_collections - 1).astype(np.int32)


def pack_mask(bits: np.ndarray) -> np.ndarray:
    """bool[n] -> uint32 words, bit r%32 of word r//32 (little-endian bit order)."""
    b = np.asarray(bits, dtype=bool)
    pad = (-len(b)) % 32
    if pad:
        b = np.concatenate([b, np.zeros(pad, dtype=bool)])
    return np.packbits(b, bitorder="little").view(np.uint32)
