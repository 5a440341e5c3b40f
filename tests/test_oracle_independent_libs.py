"""A third, library-backed statement of the leg primitives, to triangulate the oracle (which is "parity unpinned":
no qdrant-client anywhere to run it against).

The oracle (oracle/oracle.py) and its reference-shaped twin (RefShapedIndex) were both written for this repo.  Here the
SAME published semantics are computed with code nobody in this repo wrote:

  * dense leg   = cosine similarity of the stored rows with the query   -> sklearn.metrics.pairwise.cosine_similarity
                  and sklearn.neighbors.NearestNeighbors(metric="cosine", algorithm="brute") for the top-k itself;
  * sparse leg  = dot product over the common indices, a document without a common index is no candidate
                  -> scipy.sparse CSR @ CSC product; the candidate set = the structural non-zeros of that product;
  * fusion      = reciprocal rank fusion, score(d) = sum over the lists containing d of 1 / (k + rank0(d)), k = 2
                  -> pandas: concat of the ranked lists, groupby(id).sum(), stable sort.

This pins nothing about qdrant's constants (k = 2, tie order, root filter: SURVEY R4/R5/R9 stay recalled); it shows that
what the oracle computes for each primitive is what the established libraries compute for "cosine", "sparse dot" and
"RRF".  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp
from sklearn.metrics.pairwise import cosine_similarity
from sklearn.neighbors import NearestNeighbors

from helpers import Corpus
from oracle import oracle


@pytest.fixture(scope="module")
def corpus():
    return Corpus(3_000, dim=256, vocab=20_011)


def test_dense_leg_is_sklearn_cosine_topk(corpus):
    c = corpus
    x = oracle.bf16_bits_to_f32(c.bits).astype(np.float64)          # the stored rows (bf16 values)
    qf, _, _, _ = c.queries(8)
    qb = oracle.normalize_bf16(qf)
    y = oracle.bf16_bits_to_f32(qb).astype(np.float64)
    # stored rows and queries are unit vectors up to bf16 rounding: the engine's score is their DOT product, i.e. the
    # cosine of the fp32 vectors the embedder produced, up to that rounding (qdrant normalises at insert for Cosine)
    cos = cosine_similarity(x, y)                                   # [n, nq], re-normalises: differs by the bf16 norm error
    nrm = np.linalg.norm(x, axis=1)[:, None] * np.linalg.norm(y, axis=1)[None, :]
    nn = NearestNeighbors(n_neighbors=40, metric="cosine", algorithm="brute").fit(x)
    _, nbrs = nn.kneighbors(y)
    for q in range(8):
        s = oracle.dense_scores(c.bits, qb[q])
        assert np.allclose(s, cos[:, q] * nrm[:, q], atol=2e-7)     # same numbers as the library, then one cast to fp32
        ids, sc = oracle.leg_topk(s, np.ones(c.n, bool), 20)
        # the library's ranking is by cosine of RE-normalised vectors; bf16 norms are 1 +- 4e-3, so compare as sets on
        # a margin: every oracle hit must be among the library's top-40, and the two top-1 agree
        assert set(ids.tolist()) <= set(nbrs[q].tolist())
        assert ids[0] == nbrs[q][0]
        # exact ranking against the library's un-renormalised dot products (what Dot / pre-normalised Cosine computes)
        dots = (x @ y[q]).astype(np.float32)
        order = np.lexsort((np.arange(c.n), -dots.astype(np.float64)))[:20]
        assert np.array_equal(ids, order) and np.array_equal(sc, dots[order])


def test_sparse_leg_is_scipy_sparse_product(corpus):
    c = corpus
    D = sp.csr_matrix((c.w.astype(np.float64), c.terms.astype(np.int64), c.indptr), shape=(c.n, c.vocab))
    _, ip, tt, ww = c.queries(8)
    for q in range(8):
        qi, qv = tt[ip[q]:ip[q + 1]], ww[ip[q]:ip[q + 1]]
        Q = sp.csc_matrix((qv.astype(np.float64), (qi.astype(np.int64), np.zeros(len(qi), np.int64))), shape=(c.vocab, 1))
        prod = (D @ Q).tocsc()                                      # structural non-zeros = docs sharing >= 1 index
        lib_scores = np.zeros(c.n)
        lib_scores[prod.indices] = prod.data
        lib_touched = np.zeros(c.n, bool)
        lib_touched[prod.indices] = True
        s, touched = oracle.sparse_scores(c.indptr, c.terms, c.w, qi, qv)
        assert np.array_equal(touched, lib_touched)                 # the candidate set ("touched" semantics, R7)
        assert np.allclose(s, lib_scores, rtol=1e-6, atol=1e-7)     # fp64 sums of the same products, cast to fp32
        ids, sc = oracle.leg_topk(s, touched, 20)
        lib_order = np.lexsort((np.arange(c.n), -lib_scores.astype(np.float32).astype(np.float64)))
        lib_order = [i for i in lib_order if lib_touched[i]][:20]
        assert ids.tolist() == lib_order


def test_rrf_is_the_textbook_formula_pandas():
    import pandas as pd
    rng = np.random.default_rng(4)
    for trial in range(50):
        n1, n2 = rng.integers(0, 30, 2)
        a = rng.choice(60, n1, replace=False)
        b = rng.choice(60, n2, replace=False)
        df = pd.concat([pd.DataFrame({"id": a, "rr": 1.0 / (2 + np.arange(n1)), "seen": np.arange(n1)}),
                        pd.DataFrame({"id": b, "rr": 1.0 / (2 + np.arange(n2)), "seen": n1 + np.arange(n2)})])
        if len(df) == 0:
            ids, sc = oracle.rrf_fuse([a, b], 10)
            assert len(ids) == 0
            continue
        g = df.groupby("id", sort=False).agg(score=("rr", "sum"), first=("seen", "min")).reset_index()
        g = g.sort_values(["score", "first"], ascending=[False, True], kind="stable").head(10)
        ids, sc = oracle.rrf_fuse([a, b], 10)
        assert ids.tolist() == g["id"].tolist()
        assert np.allclose(sc, g["score"].to_numpy(), rtol=0, atol=1e-15)
