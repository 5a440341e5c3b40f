"""CPU oracle for the hybrid-retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
(``audio-rag_b200/b200rag``) never does and fails loudly without its CUDA library.

PARITY UNPINNED.  The arithmetic of this path does not live in the reference
repository: ``QdrantRetriever.search`` (/root/reference/src/audio_rag/retrieval/qdrant.py:227-352)
only builds a request for the third-party ``qdrant-client`` (declared ``>=1.14.3``,
pyproject.toml:25, no lockfile; server image qdrant/qdrant:v1.12.0, docker-compose.yml:31),
which is not installed here and has no wheel in /opt/wheelhouse.  The reference's own tests
hold no golden vector, known-answer test or fixture for retrieval.
This file therefore restates the *published* algorithm of qdrant-client local mode
(``qdrant_client/local/local_collection.py::search``, ``local/distances.py::cosine_similarity``,
``local/sparse_distances.py::sparse_dot_product``, ``hybrid/fusion.py::reciprocal_rank_fusion``)
as rules R1-R12 of SURVEY.md §8c, anchored on the reference's call sites:

  qdrant.py:281-298  hybrid  = prefetch[dense limit 2k, sparse limit 2k] + Fusion.RRF, limit k, root filter
  qdrant.py:299-312  sparse  = using="sparse", limit k
  qdrant.py:313-332  dense   = using="dense" | unnamed; score_threshold only on legacy collections, only if > 0
  qdrant.py:98-117   Distance.COSINE, size = embedding_dim; sparse vectors without IDF modifier
  qdrant.py:248,59   searching an unknown collection creates it empty and returns []

Two tiers live here:
  * the CANONICAL oracle (``dense_scores``/``sparse_scores``/``leg_topk``/``rrf_fuse``) which defines
    the engine's numerics bit for bit (same bf16 bits, fp64 accumulation in index order, cast to
    fp32; stated tie-break: score desc, then smaller row id);
  * the REFERENCE-SHAPED path (``RefShapedIndex``) which has the algorithmic shape of qdrant-client
    local mode (fp32 BLAS sgemv + full argsort; a Python two-pointer merge per document; dict RRF).
    It is what ``bench.py`` times as the CPU baseline ("port") and the 1e-2 comparator of R2.
"""
from __future__ import annotations

import numpy as np

RRF_K = 2  # qdrant_client/hybrid/fusion.py: ranking_constant = 2  [3P-RECALL, SURVEY R9]


# ----------------------------------------------------------------------------- bf16 / normalisation

def f32_to_bf16_bits(y):
    u = np.ascontiguousarray(y, dtype=np.float32).view(np.uint32)
    r = u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    return (r >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b):
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def normalize_f32(x):
    """R2: cosine = dot of unit vectors (local/distances.py::cosine_similarity normalises both sides).

    Defined arithmetic (the engine's host routine ``b200rag_normalize_bf16`` is the same, bit for bit):
    ss = sequential fp64 sum of squares in index order; y = fp32(fp64(x) / sqrt(ss)).
    Zero rows divide by 1 (qdrant adds EPSILON; generators never emit zero rows)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float32))
    x64 = x.astype(np.float64)
    ss = np.cumsum(x64 * x64, axis=1)[:, -1] if x.shape[1] else np.zeros(len(x))
    nrm = np.sqrt(ss)
    nrm = np.where(nrm == 0.0, 1.0, nrm)
    return (x64 / nrm[:, None]).astype(np.float32)


def normalize_bf16(x):
    return f32_to_bf16_bits(normalize_f32(x))


# ----------------------------------------------------------------------------- canonical leg scoring

def dense_scores(corpus_bits, q_bits, rows=None):
    """R2 canonical dense score, fp64 accumulate in a fixed, documented order, then one cast to fp32.

    Element k of a row belongs to lane l = (k % 256) // 8.  Lane partial p_l = sum of its products in
    ascending k (fp64 adds, sequential); score = pairwise tree over the 32 lane partials
    ((p0+p1)+(p2+p3))+... ; result fp32(score) + 0.0.  A bf16*bf16 product is exact in fp64, so fused
    or unfused multiply-add give the same bits.  (dim must be a multiple of 256.)"""
    c = np.asarray(corpus_bits, dtype=np.uint16)
    if rows is not None:
        c = c[np.asarray(rows, dtype=np.int64)]
    n, dim = c.shape
    assert dim % 256 == 0
    nch = dim // 256
    q = bf16_bits_to_f32(np.asarray(q_bits, dtype=np.uint16)).astype(np.float64)
    qv = q.reshape(nch, 32, 8).transpose(1, 0, 2).reshape(32, nch * 8)          # [lane, step]
    out = np.empty(n, dtype=np.float32)
    blk = 32768
    for s in range(0, n, blk):
        cf = bf16_bits_to_f32(c[s:s + blk]).astype(np.float64)
        m = cf.shape[0]
        cv = cf.reshape(m, nch, 32, 8).transpose(0, 2, 1, 3).reshape(m, 32, nch * 8)
        p = np.zeros((m, 32), dtype=np.float64)
        for t in range(nch * 8):
            p += cv[:, :, t] * qv[None, :, t]
        while p.shape[1] > 1:
            p = p[:, 0::2] + p[:, 1::2]
        out[s:s + blk] = p[:, 0].astype(np.float32)
    return (out + np.float32(0.0)).astype(np.float32)


def check_sparse_vector(idx, val):
    """R3: duplicate indices in one sparse vector are invalid input (qdrant rejects them)."""
    idx = np.asarray(idx, dtype=np.int64)
    val = np.asarray(val, dtype=np.float32)
    if idx.shape != val.shape:
        raise ValueError("sparse indices/values length mismatch")
    if len(np.unique(idx)) != len(idx):
        raise ValueError("duplicate index in sparse vector")
    if (idx < 0).any():
        raise ValueError("negative sparse index")
    o = np.argsort(idx, kind="stable")
    return idx[o], val[o]


def sparse_scores(indptr, terms, weights, q_idx, q_val, rows=None):
    """R3/R7 canonical: (score fp32[n], touched bool[n]).

    score = fp32( sum over common indices, ascending index, of fp64(w_q)*fp64(w_d) ), fp64 adds;
    a doc with no common index is untouched (dropped by the leg), one whose products sum to 0.0 stays."""
    indptr = np.asarray(indptr, dtype=np.int64)
    terms = np.asarray(terms)
    weights = np.asarray(weights, dtype=np.float32)
    q_idx, q_val = check_sparse_vector(q_idx, q_val)
    n = len(indptr) - 1
    rr = np.arange(n) if rows is None else np.asarray(rows, dtype=np.int64)
    acc = np.zeros(len(rr), dtype=np.float64)
    touched = np.zeros(len(rr), dtype=bool)
    if len(q_idx) == 0 or len(rr) == 0:
        return acc.astype(np.float32), touched
    # doc-major walk restricted to the requested rows; per doc terms are ascending so the
    # masked products below are visited in ascending index order, one query term at a time
    doc_of = np.repeat(np.arange(len(rr)), indptr[rr + 1] - indptr[rr])
    pos = _ranges(indptr[rr], indptr[rr + 1])
    t = terms[pos].astype(np.int64)
    w = weights[pos].astype(np.float64)
    for j in range(len(q_idx)):                       # ascending query index == ascending common index
        m = t == q_idx[j]
        d = doc_of[m]
        acc[d] = acc[d] + np.float64(q_val[j]) * w[m]
        touched[d] = True
    return (acc.astype(np.float32) + np.float32(0.0)).astype(np.float32), touched


def _ranges(starts, ends):
    lens = (ends - starts).astype(np.int64)
    tot = int(lens.sum())
    if tot == 0:
        return np.zeros(0, dtype=np.int64)
    off = np.repeat(starts - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
    return np.arange(tot, dtype=np.int64) + off


def leg_topk(scores, eligible, limit, score_threshold=None):
    """R4-R7 (local_collection.py::search walk): eligible rows only, score desc, ties -> smaller row id,
    stop at ``limit``; with a threshold drop candidates scoring below it.  Returns (ids int64, scores fp32)."""
    scores = np.asarray(scores, dtype=np.float32)
    ids = np.flatnonzero(np.asarray(eligible, dtype=bool))
    if len(ids) == 0 or limit <= 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float32)
    s = scores[ids]
    order = np.lexsort((ids, -s.astype(np.float64)))   # primary: score desc, secondary: id asc
    order = order[:limit]
    ids, s = ids[order], s[order]
    if score_threshold is not None:
        keep = s >= np.float32(score_threshold)
        ids, s = ids[keep], s[keep]
    return ids.astype(np.int64), s


def rrf_fuse(legs, limit, rrf_k=RRF_K):
    """R9/R10, hybrid/fusion.py::reciprocal_rank_fusion restated: dict accumulate 1/(k+pos) per leg in
    leg order, first sighting assigns, later sightings add (fp64); stable sort by fused score desc, so
    ties keep first-seen order in [dense leg ++ sparse leg].  Returns (ids int64, fused fp64)."""
    acc: dict[int, float] = {}
    for leg in legs:
        for i, pid in enumerate(leg):
            pid = int(pid)
            if pid in acc:
                acc[pid] += 1.0 / (rrf_k + i)
            else:
                acc[pid] = 1.0 / (rrf_k + i)
    items = sorted(acc.items(), key=lambda kv: kv[1], reverse=True)[:limit]
    return (np.asarray([k for k, _ in items], dtype=np.int64),
            np.asarray([v for _, v in items], dtype=np.float64))


# ----------------------------------------------------------------------------- an index with the boundary's semantics

class OracleIndex:
    """Row store + canonical search, rules R1-R12.  Row id = insertion order (R1)."""

    def __init__(self, dim=1024):
        self.dim = dim
        self.bits = np.zeros((0, dim), dtype=np.uint16)
        self.indptr = np.zeros(1, dtype=np.int64)
        self.terms = np.zeros(0, dtype=np.uint32)
        self.weights = np.zeros(0, dtype=np.float32)

    @property
    def n(self):
        return self.bits.shape[0]

    def add_bits(self, bits, sp_indptr=None, sp_terms=None, sp_weights=None):
        bits = np.asarray(bits, dtype=np.uint16).reshape(-1, self.dim)
        n = bits.shape[0]
        self.bits = np.concatenate([self.bits, bits])
        if sp_indptr is None:
            sp_indptr = np.zeros(n + 1, dtype=np.int64)
            sp_terms = np.zeros(0, np.uint32)
            sp_weights = np.zeros(0, np.float32)
        sp_indptr = np.asarray(sp_indptr, dtype=np.int64)
        assert len(sp_indptr) == n + 1
        self.indptr = np.concatenate([self.indptr, self.indptr[-1] + sp_indptr[1:]])
        self.terms = np.concatenate([self.terms, np.asarray(sp_terms, dtype=np.uint32)])
        self.weights = np.concatenate([self.weights, np.asarray(sp_weights, dtype=np.float32)])

    def add_f32(self, dense, sparse=None):
        """dense fp32 [n, dim]; sparse = list of (indices, values) or None per row."""
        dense = np.asarray(dense, dtype=np.float32).reshape(-1, self.dim)
        indptr = [0]
        tt, ww = [], []
        for i in range(len(dense)):
            sv = None if sparse is None else sparse[i]
            if sv is not None:
                ii, vv = check_sparse_vector(sv[0], sv[1])
                tt.append(ii.astype(np.uint32))
                ww.append(vv)
                indptr.append(indptr[-1] + len(ii))
            else:
                indptr.append(indptr[-1])
        self.add_bits(normalize_bf16(dense), np.asarray(indptr),
                      np.concatenate(tt) if tt else np.zeros(0, np.uint32),
                      np.concatenate(ww) if ww else np.zeros(0, np.float32))

    def dense_leg(self, q_bits, eligible, limit, score_threshold=None):
        return leg_topk(dense_scores(self.bits, q_bits), eligible, limit, score_threshold)

    def sparse_leg(self, q_idx, q_val, eligible, limit):
        s, touched = sparse_scores(self.indptr, self.terms, self.weights, q_idx, q_val)
        return leg_topk(s, np.asarray(eligible, dtype=bool) & touched, limit)

    def search(self, mode, q_bits, q_idx, q_val, eligible, top_k, score_threshold=None, rrf_k=RRF_K):
        """mode in {'dense','sparse','hybrid'} AFTER the R11 fallback was applied by the caller.
        Returns (ids int64, scores float64)."""
        if eligible is None:
            eligible = np.ones(self.n, dtype=bool)
        if mode == "dense":
            i, s = self.dense_leg(q_bits, eligible, top_k, score_threshold)
            return i, s.astype(np.float64)
        if mode == "sparse":
            i, s = self.sparse_leg(q_idx, q_val, eligible, top_k)
            return i, s.astype(np.float64)
        if mode == "hybrid":                                        # R8: L = 2*top_k per leg
            di, _ = self.dense_leg(q_bits, eligible, 2 * top_k)
            si, _ = self.sparse_leg(q_idx, q_val, eligible, 2 * top_k)
            return rrf_fuse([di, si], top_k, rrf_k)
        raise ValueError(mode)


# ----------------------------------------------------------------------------- reference-shaped CPU path (the timed baseline)

def sparse_dot_product_two_pointer(ai, av, bi, bv):
    """local/sparse_distances.py::sparse_dot_product shape: two-pointer merge over index-sorted vectors;
    returns None when the vectors share no index."""
    i = j = 0
    acc = 0.0
    overlap = False
    na, nb = len(ai), len(bi)
    while i < na and j < nb:
        a, b = ai[i], bi[j]
        if a == b:
            overlap = True
            acc += av[i] * bv[j]
            i += 1
            j += 1
        elif a < b:
            i += 1
        else:
            j += 1
    return acc if overlap else None


class RefShapedIndex:
    """Algorithmic shape of qdrant-client local mode (what the reference runs with
    ``qdrant_in_memory=True``, qdrant.py:42-44): fp32 unit rows + BLAS sgemv + full argsort for the
    dense leg; a Python loop over ALL documents with a two-pointer merge for the sparse leg; dict RRF."""

    def __init__(self, dense_f32, sp_indptr, sp_terms, sp_weights):
        self.vec = normalize_f32(dense_f32)
        self.indptr = np.asarray(sp_indptr, dtype=np.int64)
        self.terms = [int(x) for x in np.asarray(sp_terms)]
        self.weights = [float(x) for x in np.asarray(sp_weights)]
        self.n = self.vec.shape[0]

    def dense_leg(self, q_f32, eligible, limit, score_threshold=None):
        q = normalize_f32(q_f32)[0]
        scores = self.vec @ q                                    # sgemv
        order = np.argsort(scores)[::-1]                         # O(N log N), like local_collection.search
        out_i, out_s = [], []
        for idx in order:
            if len(out_i) >= limit:
                break
            if eligible is not None and not eligible[idx]:
                continue
            if score_threshold is not None and scores[idx] < score_threshold:
                break
            out_i.append(int(idx))
            out_s.append(float(scores[idx]))
        return out_i, out_s

    def sparse_leg(self, q_idx, q_val, eligible, limit):
        qi, qv = check_sparse_vector(q_idx, q_val)
        qi = [int(x) for x in qi]
        qv = [float(x) for x in qv]
        scores = np.full(self.n, -np.inf, dtype=np.float32)
        ip = self.indptr
        for d in range(self.n):                                  # the per-document Python loop
            s, e = ip[d], ip[d + 1]
            r = sparse_dot_product_two_pointer(qi, qv, self.terms[s:e], self.weights[s:e])
            if r is not None:
                scores[d] = r
        order = np.argsort(scores)[::-1]
        out_i, out_s = [], []
        for idx in order:
            if len(out_i) >= limit:
                break
            if scores[idx] == -np.inf:
                break
            if eligible is not None and not eligible[idx]:
                continue
            out_i.append(int(idx))
            out_s.append(float(scores[idx]))
        return out_i, out_s

    def hybrid(self, q_f32, q_idx, q_val, eligible, top_k):
        di, _ = self.dense_leg(q_f32, eligible, 2 * top_k)
        si, _ = self.sparse_leg(q_idx, q_val, eligible, 2 * top_k)
        return rrf_fuse([di, si], top_k)
