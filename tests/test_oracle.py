"""The oracle against known answers (SURVEY.md Appendix B - the facts recalled about qdrant-client, none of which
the reference's own tests pin: PARITY UNPINNED), its own invariants (hypothesis), and numpy-vs-C equivalence."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from b200rag import synth
from oracle import fast, oracle


def test_rrf_known_answers():
    # B.1: pos 0 in both legs -> 1/2 + 1/2 = 1.0 exactly; only pos 3 of one leg -> 0.2
    ids, sc = oracle.rrf_fuse([[7, 1, 2, 3], [7, 9, 8, 5]], 10)
    assert ids[0] == 7 and sc[0] == 1.0
    assert sc[list(ids).index(3)] == 1.0 / 5 and sc[list(ids).index(5)] == 0.2
    # B.2: dense-only candidate at pos i ties with the sparse-only one at pos i -> dense first (stable sort)
    ids, sc = oracle.rrf_fuse([[10, 11], [20, 21]], 4)
    assert list(ids) == [10, 20, 11, 21] and sc[0] == sc[1] == 0.5 and sc[2] == sc[3] == 1.0 / 3
    # B.5: top_k = 1 needs the depth-2 prefetch: #2 in both legs beats each leg's #1
    ids, sc = oracle.rrf_fuse([[1, 5], [2, 5]], 1)
    assert list(ids) == [5] and sc[0] == 2.0 / 3
    # fused scores are plain fp64 sums in leg order
    ids, sc = oracle.rrf_fuse([[1, 2, 3], [3, 1]], 3)
    assert sc[list(ids).index(1)] == 1.0 / 2 + 1.0 / 3 and sc[list(ids).index(3)] == 1.0 / 4 + 1.0 / 2
    # empty legs
    ids, sc = oracle.rrf_fuse([[], []], 5)
    assert len(ids) == 0


def test_sparse_touched_semantics():
    # B.3: no-overlap docs are absent; a doc overlapping only on a zero weight is present with score 0.0
    indptr = np.array([0, 2, 3, 3, 5])
    terms = np.array([1, 4, 4, 2, 9], np.uint32)
    w = np.array([0.5, 2.0, 0.0, 1.0, 1.0], np.float32)
    s, touched = oracle.sparse_scores(indptr, terms, w, [4, 1], [3.0, 2.0])
    assert list(touched) == [True, True, False, False]
    assert s[0] == np.float32(0.5 * 2.0 + 2.0 * 3.0) and s[1] == 0.0
    ids, sc = oracle.leg_topk(s, touched, 10)
    assert list(ids) == [0, 1]
    with pytest.raises(ValueError):
        oracle.sparse_scores(indptr, terms, w, [4, 4], [1.0, 1.0])
    s2, t2 = fast.sparse_scores(indptr, terms, w, [4, 1], [3.0, 2.0])
    assert np.array_equal(s, s2) and np.array_equal(touched, t2)


def test_cosine_is_scale_invariant_and_tie_break_is_row_order():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((50, 256)).astype(np.float32)
    x[7] = x[3]            # exact duplicates: R5 -> smaller row id first
    x[30] = x[3] * 5.0     # B.4: un-normalised copy, same direction
    oi = oracle.OracleIndex(256)
    oi.add_f32(x)
    q = oracle.normalize_bf16(x[3:4] * 0.01)[0]
    ids, sc = oi.dense_leg(q, np.ones(50, bool), 3)
    assert list(ids) == [3, 7, 30] and sc[0] == sc[1] == sc[2]
    # B.7: threshold drops low scores
    ids2, _ = oi.dense_leg(q, np.ones(50, bool), 50, score_threshold=0.9)
    assert list(ids2) == [3, 7, 30]
    # masks remove rows BEFORE selection (R4)
    elig = np.ones(50, bool)
    elig[3] = False
    ids3, _ = oi.dense_leg(q, elig, 2)
    assert list(ids3) == [7, 30]


def test_normalize_matches_library_host_routine(built_lib):
    from b200rag import normalize_bf16
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((300, 1024)) * rng.uniform(0.01, 100, (300, 1))).astype(np.float32)
    x[5] = 0.0
    assert np.array_equal(normalize_bf16(x), oracle.normalize_bf16(x))
    with pytest.raises(Exception):
        normalize_bf16(np.array([[np.nan] * 256], np.float32))


@pytest.mark.parametrize("dim", [256, 768, 1024])
def test_numpy_and_c_oracles_agree(dim):
    bits = fast.synth_dense_bf16(11, 0, 700, dim)
    assert np.array_equal(bits, synth.dense_rows_bf16(11, 0, 700, dim))
    q = oracle.normalize_bf16(synth.dense_queries_f32(12, 0, 2, 700, dim, corpus_seed=11))
    for i in range(2):
        a, b = oracle.dense_scores(bits, q[i]), fast.dense_scores(bits, q[i])
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # the 1e-2 comparator of R2: fp32 BLAS on the unrounded vectors stays close to the canonical bf16 score
    f = synth.dense_rows_f32(11, 0, 700, dim)
    qf = synth.dense_queries_f32(12, 0, 1, 700, dim, corpus_seed=11)[0]
    canon = oracle.dense_scores(bits, oracle.normalize_bf16(qf[None])[0])
    top = np.argsort(-canon)[:10]
    assert np.allclose((f @ qf)[top], canon[top], rtol=1e-2, atol=2e-3)
    assert abs((f @ qf)[top[0]] - canon[top[0]]) <= 1e-2 * abs(canon[top[0]])


def test_sparse_numpy_and_c_agree_on_zipf_corpus():
    thr = synth.zipf_thresholds(5003)
    idf, tff = synth.bm25_tables(4000, 5003)
    ip, tt, ww = synth.sparse_docs_csr(5, 0, 4000, 4000, 5003, 256, thr, (idf, tff))
    ip2, tt2, ww2 = fast.synth_sparse_csr(5, 0, 4000, thr, idf, tff, 5003, 256, synth.TERM_PERM_MUL % 5003)
    assert np.array_equal(ip, ip2) and np.array_equal(tt, tt2) and np.array_equal(ww.view(np.uint32), ww2.view(np.uint32))
    qi, qt, qw = synth.sparse_queries(6, 0, 5, 12, 5003, thr)
    for i in range(5):
        a, ta = oracle.sparse_scores(ip, tt, ww, qt[qi[i]:qi[i + 1]], qw[qi[i]:qi[i + 1]])
        b, tb = fast.sparse_scores(ip, tt, ww, qt[qi[i]:qi[i + 1]], qw[qi[i]:qi[i + 1]])
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ta, tb)


def test_ref_shaped_path_agrees_with_canonical_on_ids():
    """The qdrant-local-shaped CPU path (fp32 sgemv + Python two-pointer loop + dict RRF) returns the same ids as the
    canonical oracle on planted queries (scores differ only by bf16 rounding, <= 1e-2 relative)."""
    n, dim, V = 600, 256, 3001
    f = synth.dense_rows_f32(3, 0, n, dim)
    thr = synth.zipf_thresholds(V)
    ip, tt, ww = synth.sparse_docs_csr(3, 0, n, n, V, 64, thr, synth.bm25_tables(n, V, 64))
    ref = oracle.RefShapedIndex(f, ip, tt, ww)
    oi = oracle.OracleIndex(dim)
    oi.add_bits(oracle.normalize_bf16(f), ip, tt, ww)
    qf = synth.dense_queries_f32(4, 0, 4, n, dim, corpus_seed=3)
    qi, qt, qw = synth.sparse_queries(4, 0, 4, 8, V, thr)
    for i in range(4):
        if i % 10 == 9:
            continue
        sl = slice(qi[i], qi[i + 1])
        di, ds = ref.dense_leg(qf[i], None, 5)
        ci, cs = oi.dense_leg(oracle.normalize_bf16(qf[i:i + 1])[0], np.ones(n, bool), 5)
        assert di[0] == ci[0] and np.allclose(ds, cs[:len(ds)], rtol=1e-2, atol=1e-3)
        si, ss = ref.sparse_leg(qt[sl], qw[sl], None, 5)
        ki, ks = oi.sparse_leg(qt[sl], qw[sl], np.ones(n, bool), 5)
        assert np.allclose(ss, ks, rtol=1e-5)
        hi, hs = ref.hybrid(qf[i], qt[sl], qw[sl], None, 3)
        assert len(hi) == 3


ids_st = st.lists(st.integers(0, 40), min_size=0, max_size=12, unique=True)


@settings(max_examples=200, deadline=None)
@given(ids_st, ids_st, st.integers(1, 15))
def test_rrf_properties(dense, sparse, k):
    ids, sc = oracle.rrf_fuse([dense, sparse], k)
    union = list(dict.fromkeys(dense + sparse))
    assert len(ids) == min(k, len(union))
    assert len(set(ids.tolist())) == len(ids)
    assert all(sc[i] >= sc[i + 1] for i in range(len(sc) - 1))
    for pid, s in zip(ids, sc):
        exp = 0.0
        first = True
        for leg in (dense, sparse):
            if pid in leg:
                v = 1.0 / (2 + leg.index(pid))
                exp = v if first else exp + v
                first = False
        assert s == exp
    # ties keep first-seen order in [dense ++ sparse]
    order = {p: i for i, p in enumerate(union)}
    for a in range(len(ids) - 1):
        if sc[a] == sc[a + 1]:
            assert order[int(ids[a])] < order[int(ids[a + 1])]


@settings(max_examples=100, deadline=None)
@given(st.lists(st.floats(-1, 1, width=32), min_size=1, max_size=60), st.integers(1, 20), st.data())
def test_leg_topk_properties(scores, limit, data):
    s = np.asarray(scores, np.float32) + np.float32(0.0)
    elig = np.asarray(data.draw(st.lists(st.booleans(), min_size=len(s), max_size=len(s))))
    ids, sc = oracle.leg_topk(s, elig, limit)
    assert len(ids) == min(limit, int(elig.sum()))
    assert all(elig[i] for i in ids)
    for a in range(len(ids) - 1):
        assert sc[a] > sc[a + 1] or (sc[a] == sc[a + 1] and ids[a] < ids[a + 1])
    if len(ids):
        rest = [i for i in np.flatnonzero(elig) if i not in set(ids.tolist())]
        assert all(s[i] <= sc[-1] for i in rest)
