"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same
seeded inputs.  Bar: chunk ids bit-exact under the stated tie-break (score desc, smaller row id first; RRF ties
keep first-seen order), dense/sparse leg scores bit-equal to the oracle's fp32(fp64 canonical sum), RRF
scores bit-equal fp64.  (The north star allows 1e-5 relative for fp32 accumulate; the engine re-scores its
candidates in the oracle's own operation order, so the tolerance used here is 0.)"""
import numpy as np
import pytest

from helpers import Corpus, assert_result_equal, oracle_search

pytestmark = pytest.mark.gpu


def _shard_from(c, gpu, **kw):
    from b200rag import Shard
    sh = Shard(dim=c.dim, vocab=c.vocab, device=gpu, row_base=kw.pop("row_base", 0), **kw)
    sh.add(c.bits, c.indptr, c.terms, c.w)
    return sh


def _run_and_check(sh, c, mode, nq, top_k, masks=None, mask_ids=None, score_threshold=None, qid_start=0,
                   n_tokens=12, batch=None):
    from b200rag import normalize_bf16
    qf, ip, tt, ww = c.queries(nq, qid_start=qid_start, n_tokens=n_tokens)
    qb = normalize_bf16(qf)
    batch = batch or nq
    for s in range(0, nq, batch):
        e = min(nq, s + batch)
        sub_ip = ip[s:e + 1] - ip[s]
        sub_t, sub_w = tt[ip[s]:ip[e]], ww[ip[s]:ip[e]]
        mids = None if mask_ids is None else np.asarray(mask_ids[s:e], dtype=np.int32)
        ids, scores, counts = sh.search(mode, top_k, qb[s:e], sub_ip, sub_t, sub_w, mask_ids=mids,
                                        score_threshold=score_threshold)
        for i in range(s, e):
            elig = None
            if mids is not None and mids[i - s] >= 0:
                elig = masks[int(mids[i - s])]
            ei, es = oracle_search(c, mode, qb[i], tt[ip[i]:ip[i + 1]], ww[ip[i]:ip[i + 1]], elig, top_k,
                                   score_threshold, row_base=sh.row_base)
            assert_result_equal(ids[i - s], scores[i - s], int(counts[i - s]), ei, es,
                                ctx=f"mode={mode} q={i} k={top_k} n={c.n}")


def test_device_generators_match_host_twins(gpu):
    import torch
    from b200rag import Shard, synth
    c = Corpus(3000, dim=1024, n_total=50_000, row_start=777)
    sh = Shard(dim=1024, vocab=c.vocab, device=gpu)
    dev = torch.device("cuda", gpu)
    out = torch.empty((c.n, 1024), dtype=torch.int16, device=dev)
    sh.synth_dense(c.seed, 777, c.n, out)
    sh.sync()
    assert np.array_equal(out.cpu().numpy().view(np.uint16), c.bits)
    thr = torch.from_numpy(c.thr.view(np.int64)).to(dev)
    idf = torch.from_numpy(c.idf).to(dev)
    tff = torch.from_numpy(c.tff).to(dev)
    counts = torch.empty(c.n, dtype=torch.int64, device=dev)
    mul = synth.TERM_PERM_MUL % c.vocab
    sh.synth_sparse(c.seed, 777, c.n, 256, thr, idf, tff, mul, counts, None, None, None)
    indptr = torch.empty(c.n + 1, dtype=torch.int64, device=dev)
    sh.exclusive_scan_i64(counts, c.n, indptr)
    sh.sync()
    assert np.array_equal(indptr.cpu().numpy(), c.indptr)
    nnz = int(indptr[-1].item())
    terms = torch.empty(nnz, dtype=torch.int32, device=dev)
    w = torch.empty(nnz, dtype=torch.float32, device=dev)
    sh.synth_sparse(c.seed, 777, c.n, 256, thr, idf, tff, mul, None, indptr, terms, w)
    sh.sync()
    assert np.array_equal(terms.cpu().numpy().view(np.uint32), c.terms)
    assert np.array_equal(w.cpu().numpy().view(np.uint32), c.w.view(np.uint32))
    # collection masks
    cthr = synth.zipf_thresholds(1000)
    cthr_d = torch.from_numpy(cthr.view(np.int64)).to(dev)
    words = torch.zeros((c.n + 31) // 32, dtype=torch.int32, device=dev)
    sh.synth_collection_mask(c.seed, 777, c.n, cthr_d, 1000, 0, words)
    sh.sync()
    exp = synth.pack_mask(synth.row_collections(c.seed, 777, c.n, 1000, cthr) == 0)
    assert np.array_equal(words.cpu().numpy().view(np.uint32), exp)
    # read-back of stored rows is the identity
    sh.add(c.bits[:100])
    assert np.array_equal(sh.read_dense(0, 100), c.bits[:100])
    sh.close()


@pytest.mark.parametrize("dim", [256, 768, 1024])
@pytest.mark.parametrize("nq,batch", [(4, 1), (4, 2), (3, 3)])
def test_dense_parity(gpu, dim, nq, batch):
    c = Corpus(20_011, dim=dim, sparse=False)
    sh = _shard_from(c, gpu)
    _run_and_check(sh, c, "dense", nq, 10, batch=batch)
    sh.close()


@pytest.mark.parametrize("top_k", [1, 5, 100])
def test_dense_topk_sizes(gpu, top_k):
    c = Corpus(9_973, dim=1024, sparse=False)
    sh = _shard_from(c, gpu)
    _run_and_check(sh, c, "dense", 3, top_k, batch=1)
    sh.close()


@pytest.mark.parametrize("R", [1024, 8192])
def test_sparse_parity(gpu, R):
    c = Corpus(20_011, dim=256, vocab=30_011)
    sh = _shard_from(c, gpu, docs_per_block=R)
    _run_and_check(sh, c, "sparse", 6, 10, batch=3)
    _run_and_check(sh, c, "sparse", 2, 100, batch=2, qid_start=50)
    sh.close()


def test_sparse_long_queries_and_full_vocab(gpu):
    c = Corpus(6_000, dim=256)              # V = 250 002
    sh = _shard_from(c, gpu, docs_per_block=2048)
    _run_and_check(sh, c, "sparse", 2, 10, n_tokens=12)
    _run_and_check(sh, c, "sparse", 2, 10, n_tokens=256 + 200, qid_start=7)   # HyDE-sized query, > 1 term chunk
    sh.close()


def test_hybrid_parity_config1(gpu):
    """BASELINE config 1: hybrid, 10k chunks, single query, top_k = 5."""
    c = Corpus(10_000, dim=1024)
    sh = _shard_from(c, gpu)
    _run_and_check(sh, c, "hybrid", 8, 5, batch=1)
    _run_and_check(sh, c, "hybrid", 4, 10, batch=4, qid_start=100)
    sh.close()


def test_masks_both_legs(gpu):
    from b200rag import synth
    c = Corpus(12_345, dim=256, vocab=30_011)
    sh = _shard_from(c, gpu, docs_per_block=4096)
    coll = synth.row_collections(c.seed, 0, c.n, 7)
    masks = {}
    for m in range(7):
        masks[m] = coll == m
        sh.mask_set(m, synth.pack_mask(masks[m]), c.n)
    mask_ids = [0, 3, -1, 6, 1, 2]
    for mode in ("dense", "sparse", "hybrid"):
        _run_and_check(sh, c, mode, 6, 10, masks=masks, mask_ids=mask_ids, batch=3)
    # a mask with fewer eligible rows than top_k, and an empty one
    small = np.zeros(c.n, dtype=bool)
    small[[5, 77, 4000]] = True
    masks[10] = small
    masks[11] = np.zeros(c.n, dtype=bool)
    sh.mask_set(10, synth.pack_mask(small), c.n)
    sh.mask_set(11, synth.pack_mask(masks[11]), c.n)
    for mode in ("dense", "sparse", "hybrid"):
        _run_and_check(sh, c, mode, 2, 10, masks=masks, mask_ids=[10, 11], batch=2)
    sh.close()


def test_edge_cases(gpu):
    from b200rag import Shard, normalize_bf16
    from oracle import oracle
    dim = 256
    sh = Shard(dim=dim, vocab=1000, device=gpu, docs_per_block=1024)
    q = normalize_bf16(np.ones((1, dim), np.float32))
    sp = (np.array([0, 2]), np.array([3, 9], np.uint32), np.array([1.0, 2.0], np.float32))
    # empty shard (R11: unknown/empty collection -> [])
    for mode in ("dense", "sparse", "hybrid"):
        ids, scores, counts = sh.search(mode, 5, q, *sp)
        assert counts[0] == 0
    # duplicate rows: exact ties resolve to the smaller row id; more duplicates than top_k + slack
    rng = np.random.default_rng(0)
    base = rng.standard_normal((40, dim)).astype(np.float32)
    rows = np.concatenate([base, np.repeat(base[:1], 80, axis=0), base[1:3]])
    bits = normalize_bf16(rows)
    # sparse: rows 0..9 share term 3 with weight 0 (touched, score 0.0), row 10 has term 9, others nothing
    indptr = [0]
    terms, w = [], []
    for r in range(len(rows)):
        if r < 10:
            terms.append(3); w.append(0.0)
        elif r == 10:
            terms += [3, 9]; w += [0.5, 1.5]
        indptr.append(len(terms))
    sh.add(bits, np.array(indptr), np.array(terms, np.uint32), np.array(w, np.float32))
    oi = oracle.OracleIndex(dim)
    oi.add_bits(bits, np.array(indptr), np.array(terms, np.uint32), np.array(w, np.float32))
    qq = normalize_bf16(base[:1])
    for mode, k in (("dense", 10), ("dense", 100), ("sparse", 5), ("sparse", 20), ("hybrid", 5), ("hybrid", 30)):
        ids, scores, counts = sh.search(mode, k, qq, *sp)
        ei, es = oi.search(mode, qq[0], sp[1], sp[2], None, k)
        assert_result_equal(ids[0], scores[0], int(counts[0]), ei, es, ctx=f"edge {mode} k={k}")
    st = sh.stats()
    # score_threshold (dense legacy collections, qdrant.py:331)
    ids, scores, counts = sh.search("dense", 100, qq, score_threshold=0.9)
    ei, es = oi.search("dense", qq[0], None, None, None, 100, score_threshold=0.9)
    assert_result_equal(ids[0], scores[0], int(counts[0]), ei, es, ctx="threshold")
    assert counts[0] == 81
    # query with no sparse terms in a hybrid batch (ragged): sparse leg empty -> RRF over the dense leg alone
    sp2 = (np.array([0, 0, 2]), sp[1], sp[2])
    q2 = np.concatenate([qq, normalize_bf16(base[5:6])])
    ids, scores, counts = sh.search("hybrid", 5, q2, *sp2)
    ei, es = oi.search("hybrid", q2[0], np.zeros(0, np.int64), np.zeros(0, np.float32), None, 5)
    assert_result_equal(ids[0], scores[0], int(counts[0]), ei, es, ctx="ragged q0")
    ei, es = oi.search("hybrid", q2[1], sp[1], sp[2], None, 5)
    assert_result_equal(ids[1], scores[1], int(counts[1]), ei, es, ctx="ragged q1")
    # invalid inputs fail loudly
    from b200rag import B200RagError
    with pytest.raises(B200RagError):
        sh.search("sparse", 5, qq, np.array([0, 2]), np.array([9, 3], np.uint32), np.array([1, 1], np.float32))
    with pytest.raises(B200RagError):
        sh.search("sparse", 5, qq, np.array([0, 1]), np.array([1000], np.uint32), np.array([1], np.float32))
    with pytest.raises(B200RagError):
        sh.search("dense", 0, qq)
    assert st["kernel_launches"] > 0
    sh.close()


def test_incremental_add_and_rebuild(gpu):
    c = Corpus(9_000, dim=256, vocab=30_011)
    from b200rag import Shard
    sh = Shard(dim=256, vocab=c.vocab, device=gpu, docs_per_block=2048)
    cuts = [0, 1500, 1501, 5000, 9000]
    for a, b in zip(cuts[:-1], cuts[1:]):
        ip = c.indptr[a:b + 1] - c.indptr[a]
        sh.add(c.bits[a:b], ip, c.terms[c.indptr[a]:c.indptr[b]], c.w[c.indptr[a]:c.indptr[b]])
        sub = Corpus.__new__(Corpus)
        sub.__dict__.update(c.__dict__)
        sub.n = b
        sub.bits, sub.indptr = c.bits[:b], c.indptr[:b + 1]
        sub.terms, sub.w = c.terms[:c.indptr[b]], c.w[:c.indptr[b]]
        _run_and_check(sh, sub, "hybrid", 2, 5, batch=2)
    assert sh.count == 9000 and sh.postings == len(c.terms)
    sh.clear()
    assert sh.count == 0
    sh.close()


def test_two_shards_fuse_equals_one(gpu):
    """Row-sharded composition (stage -> legs per shard -> concatenate -> fuse) == single shard, bit for bit."""
    import torch
    from b200rag import Shard, normalize_bf16
    from b200rag._ffi import CAND_DTYPE
    c = Corpus(16_000, dim=256, vocab=30_011)
    cut = 6_500
    one = _shard_from(c, gpu, docs_per_block=2048)
    a = Shard(dim=256, vocab=c.vocab, device=gpu, row_base=0, docs_per_block=2048)
    b = Shard(dim=256, vocab=c.vocab, device=gpu, row_base=cut, docs_per_block=2048)
    a.add(c.bits[:cut], c.indptr[:cut + 1], c.terms[:c.indptr[cut]], c.w[:c.indptr[cut]])
    b.add(c.bits[cut:], c.indptr[cut:] - c.indptr[cut], c.terms[c.indptr[cut]:], c.w[c.indptr[cut]:])
    qf, ip, tt, ww = c.queries(5)
    qb = normalize_bf16(qf)
    dev = torch.device("cuda", gpu)
    # top-10: gathered sets of <= 128 candidates are ranked by counting; top-100: merged by rank (sorted shard blocks)
    for mode, k in (("dense", 10), ("sparse", 10), ("hybrid", 10), ("dense", 100), ("sparse", 100), ("hybrid", 100)):
        ids1, sc1, cnt1 = one.search(mode, k, qb, ip, tt, ww)
        q, keep = a.make_query(mode, k, qb, ip, tt, ww)
        nlegs, L = Shard.legs_len(q)
        gathered = torch.zeros((2, nlegs, 5, L, 2), dtype=torch.int64, device=dev)
        amb = torch.zeros(1, dtype=torch.int32, device=dev)
        for r, sh in enumerate((a, b)):
            sh.stage(q, keep)
            sh.legs(gathered[r], amb)
            sh.sync()
        out_ids = torch.empty((5, k), dtype=torch.int64, device=dev)
        out_sc = torch.empty((5, k), dtype=torch.float64, device=dev)
        out_cnt = torch.empty(5, dtype=torch.int32, device=dev)
        a.fuse(gathered, 2, out_ids, out_sc, out_cnt)
        a.sync()
        assert int(amb.item()) == 0
        assert np.array_equal(out_cnt.cpu().numpy(), cnt1)
        assert np.array_equal(out_ids.cpu().numpy(), ids1)
        assert np.array_equal(out_sc.cpu().numpy(), sc1)
        g = gathered.cpu().numpy().view(CAND_DTYPE)
        assert g["valid"].max() == 1
        if k == 100:
            # blocks that are NOT in leg order (a foreign caller): the kernel notices and sorts instead of merging
            gen = torch.Generator(device="cpu").manual_seed(7)
            shuffled = gathered.clone()
            for r in range(2):
                for leg in range(nlegs):
                    for qq in range(5):
                        perm = torch.randperm(L, generator=gen).to(dev)
                        shuffled[r, leg, qq] = gathered[r, leg, qq][perm]
            o2, s2, c2 = torch.empty_like(out_ids), torch.empty_like(out_sc), torch.empty_like(out_cnt)
            a.fuse(shuffled, 2, o2, s2, c2)
            a.sync()
            assert torch.equal(o2, out_ids) and torch.equal(s2, out_sc) and torch.equal(c2, out_cnt)
    for sh in (one, a, b):
        sh.close()


# ------------------------------------------------------------------------------------------------ tcgen05 GEMM path
@pytest.mark.parametrize("dim,n,B", [(1024, 1000, 5), (256, 777, 3), (768, 5000, 130), (1024, 40_000, 64)])
def test_gemm_raw_scores_match_fp32_matmul(gpu, dim, n, B):
    """The tensor-core path as a GEMM: raw fp32-accumulated scores of every (query, row) pair against numpy on the
    same bf16 bits (fp32 accumulate tolerance: 1e-5 absolute on unit vectors, the north star's fp32 bound)."""
    import torch
    from b200rag import Shard, normalize_bf16, synth
    c = Corpus(n, dim=dim, sparse=False)
    sh = Shard(dim=dim, vocab=1000, device=gpu)
    sh.add(c.bits)
    qf, _, _, _ = c.queries(B)
    qb = normalize_bf16(qf)
    q, keep = sh.make_query("dense", 10, qb)
    sh.stage(q, keep)
    out = torch.full((B, n), float("nan"), dtype=torch.float32, device=torch.device("cuda", gpu))
    sh.debug_dense_scores(out)
    got = out.cpu().numpy()
    exp = synth.bf16_bits_to_f32(qb).astype(np.float64) @ synth.bf16_bits_to_f32(c.bits).astype(np.float64).T
    assert not np.isnan(got).any()
    err = np.abs(got - exp).max()
    assert err < 1e-5, f"max abs err {err}"
    sh.close()


@pytest.mark.parametrize("B,top_k", [(1, 10), (3, 10), (64, 10), (130, 5), (7, 100), (40, 20)])
def test_dense_parity_gemm_path(gpu, B, top_k):
    c = Corpus(30_011, dim=1024, sparse=False)
    sh = _shard_from(c, gpu)
    sh.set_dense_path(2)
    _run_and_check(sh, c, "dense", B, top_k, batch=B)
    assert sh.stats()["dense_path"] == 2
    sh.close()


def test_hybrid_masks_gemm_path(gpu):
    from b200rag import synth
    c = Corpus(12_345, dim=256, vocab=30_011)
    sh = _shard_from(c, gpu, docs_per_block=4096)
    sh.set_dense_path(2)
    coll = synth.row_collections(c.seed, 0, c.n, 5)
    masks = {m: coll == m for m in range(5)}
    for m in range(5):
        sh.mask_set(m, synth.pack_mask(masks[m]), c.n)
    mask_ids = [0, 3, -1, 4, 1, 2, 2, 0]
    _run_and_check(sh, c, "hybrid", 8, 10, masks=masks, mask_ids=mask_ids, batch=8)
    _run_and_check(sh, c, "dense", 8, 10, masks=masks, mask_ids=mask_ids, batch=4)
    sh.close()


def test_large_topk_and_batches_all_paths(gpu):
    """top-100 hybrid (Lc = 300: shared-memory candidate lists, several query passes), a batch above 128 queries, masks."""
    from b200rag import synth
    c = Corpus(25_000, dim=256, vocab=40_009)
    sh = _shard_from(c, gpu, docs_per_block=4096)
    coll = synth.row_collections(c.seed, 0, c.n, 3)
    masks = {m: coll == m for m in range(3)}
    for m in range(3):
        sh.mask_set(m, synth.pack_mask(masks[m]), c.n)
    for path in (0, 2):
        sh.set_dense_path(path)
        _run_and_check(sh, c, "hybrid", 5, 100, masks=masks, mask_ids=[0, 1, 2, -1, 0], batch=5)
        _run_and_check(sh, c, "dense", 40, 100, batch=40, qid_start=10)
    sh.set_dense_path(0)
    _run_and_check(sh, c, "hybrid", 150, 10, batch=150, qid_start=300)
    _run_and_check(sh, c, "sparse", 70, 20, batch=70, qid_start=500)
    sh.close()


def test_gemm_filter_path_retry_and_robust_lists(gpu):
    """tcgen05 path, top-100 with a query batch: (a) the sample+filter epilogue, (b) the shared-memory list epilogue it
    replaces (forced through a fixed slack), (c) a corpus built so that the 1/16 sample over-estimates the threshold
    (the sampled tiles hold 40 near-copies of every query, the rest of the shard none): the filter pass collects fewer
    than Lc rows, the search must notice (ambiguous) and repeat on the robust path with the exact answer."""
    from b200rag import normalize_bf16, synth
    c = Corpus(20_000, dim=256, sparse=False)
    qf, _, _, _ = c.queries(6)
    qb = normalize_bf16(qf)
    # rows 0..239 (tile 0, always in the sample) become the queries' nearest neighbours: 40 perturbed copies of each
    rng = np.random.default_rng(5)
    planted = np.repeat(qf, 40, axis=0) + 0.02 * rng.standard_normal((240, 256)).astype(np.float32)
    c.bits[:240] = normalize_bf16(planted)
    sh = _shard_from(c, gpu)
    sh.set_dense_path(2)

    def check(tag):
        ids, scores, counts = sh.search("dense", 100, qb)
        for i in range(6):
            ei, es = oracle_search(c, "dense", qb[i], None, None, None, 100)
            assert_result_equal(ids[i], scores[i], int(counts[i]), ei, es, ctx=f"{tag} q={i}")
        return sh.stats()

    st = check("filter+retry")
    assert st["retries"] >= 1, "the planted sample must force the robust retry"
    sh.set_slack(60)                      # fixed slack -> list epilogue from the start
    st = check("robust lists")
    assert st["retries"] == 0
    sh.set_slack(0)
    sh.close()


@pytest.mark.parametrize("R", [1024, 16384])
def test_sparse_fixed_point_edge_cases(gpu, R):
    """The sparse scan accumulates in fixed point with integer atomics: mixed-sign weights (documents AND queries),
    weights spanning 6 orders of magnitude, zero weights (touched, score 0), exact score ties across thousands of
    documents (bisection fallback), long queries (count bits), several blocks per CTA -- ids and scores must stay
    bit-equal to the oracle."""
    from b200rag import Shard
    rng = np.random.default_rng(R)
    n, vocab, dim = 6000, 400, 256
    c = Corpus(n, dim=dim, sparse=False)
    nnz_per = rng.integers(0, 40, size=n)
    indptr = np.concatenate([[0], np.cumsum(nnz_per)]).astype(np.int64)
    terms = np.concatenate([np.sort(rng.choice(vocab, size=k, replace=False)) for k in nnz_per]).astype(np.uint32)
    mag = 10.0 ** rng.uniform(-3, 3, size=len(terms))
    w = (mag * rng.choice([-1.0, 1.0], size=len(terms))).astype(np.float32)
    w[rng.random(len(terms)) < 0.05] = 0.0
    # documents 100..3099 carry the SAME single posting: thousands of exact ties
    tie_rows = np.arange(100, 3100)
    for r in tie_rows:
        s, e = indptr[r], indptr[r + 1]
        if e > s:
            terms[s:e] = np.sort(rng.choice(np.arange(1, vocab), size=e - s, replace=False))
            terms[s], w[s] = 0, 0.75
            w[s + 1:e] = 0.0
    c.indptr, c.terms, c.w, c.vocab = indptr, terms, w, vocab
    sh = Shard(dim=dim, vocab=vocab, device=gpu, docs_per_block=R)
    sh.add(c.bits, indptr, terms, w)
    qb = __import__("b200rag").normalize_bf16(c.queries(1)[0])
    cases = [
        (np.array([0], np.uint32), np.array([2.0], np.float32)),                                  # the tie term alone
        (np.arange(0, vocab, 3, dtype=np.uint32), rng.normal(size=len(range(0, vocab, 3))).astype(np.float32)),
        (np.arange(vocab, dtype=np.uint32), (10.0 ** rng.uniform(-2, 2, size=vocab)).astype(np.float32)),  # 400 terms
        (np.array([5, 17, 200], np.uint32), np.array([-3.0, 0.0, 1e-3], np.float32)),
    ]
    for qi, (qt, qw) in enumerate(cases):
        for k in (5, 100):
            ids, scores, counts = sh.search("sparse", k, None, np.array([0, len(qt)], np.int64), qt, qw)
            ei, es = oracle_search(c, "sparse", None, qt, qw, None, k)
            assert_result_equal(ids[0], scores[0], int(counts[0]), ei, es, ctx=f"R={R} case {qi} k={k}")
    # a batch mixing all of them (different scales per query in one launch) + a hybrid with the dense leg
    ip = np.concatenate([[0], np.cumsum([len(t) for t, _ in cases])]).astype(np.int64)
    tt = np.concatenate([t for t, _ in cases]); ww = np.concatenate([x for _, x in cases])
    ids, scores, counts = sh.search("sparse", 20, None, ip, tt, ww)
    for qi, (qt, qw) in enumerate(cases):
        ei, es = oracle_search(c, "sparse", None, qt, qw, None, 20)
        assert_result_equal(ids[qi], scores[qi], int(counts[qi]), ei, es, ctx=f"R={R} batch case {qi}")
    ids, scores, counts = sh.search("hybrid", 10, qb, np.array([0, len(cases[1][0])], np.int64), cases[1][0], cases[1][1])
    ei, es = oracle_search(c, "hybrid", qb[0], cases[1][0], cases[1][1], None, 10)
    assert_result_equal(ids[0], scores[0], int(counts[0]), ei, es, ctx=f"R={R} hybrid")
    sh.close()


def test_device_resident_queries(gpu):
    """queries that never leave the GPU.  normalize_bf16_device is bit-equal to the host routine;
    stage_device + legs + fuse returns exactly what the host-buffer search returns; bad device-side sparse input is
    refused; several staged slots can be replayed in any order."""
    import torch
    from b200rag import Shard, normalize_bf16
    from b200rag._ffi import B200RagError
    c = Corpus(9_000, dim=1024, vocab=30_011)
    sh = _shard_from(c, gpu, docs_per_block=2048)
    dev = torch.device("cuda", gpu)
    sh.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    B, k = 5, 10
    qf, ip, tt, ww = c.queries(B)
    x = torch.from_numpy(np.ascontiguousarray(qf, dtype=np.float32)).to(dev)
    bits = torch.empty((B, 1024), dtype=torch.int16, device=dev)
    sh.normalize_bf16_device(x, B, bits)
    qb = normalize_bf16(qf)
    assert np.array_equal(bits.cpu().numpy().view(np.uint16), qb)
    t_dev = torch.from_numpy(tt.astype(np.int32)).to(dev)
    w_dev = torch.from_numpy(ww).to(dev)
    # (slot 0 is the one the host-buffer search below re-stages: the device batches live in slots 1..3)
    for slot, mode in enumerate(("hybrid", "dense", "sparse"), start=1):
        sh.stage_device(mode, k, B, bits if mode != "sparse" else None, ip if mode != "dense" else None,
                        t_dev if mode != "dense" else None, w_dev if mode != "dense" else None, slot=slot)
    for slot, mode in reversed(list(enumerate(("hybrid", "dense", "sparse"), start=1))):
        sh.use_slot(slot)
        nl = 2 if mode == "hybrid" else 1
        L = 2 * k if mode == "hybrid" else k
        cands = torch.zeros((nl * B * L + 1, 2), dtype=torch.int64, device=dev)
        oi = torch.empty((B, k), dtype=torch.int64, device=dev)
        osc = torch.empty((B, k), dtype=torch.float64, device=dev)
        oc = torch.empty(B + 1, dtype=torch.int32, device=dev)
        sh.legs(cands, cands[-1])
        sh.fuse(cands, 1, oi, osc, oc, has_trailer=True)
        sh.sync()
        ids, scores, counts = sh.search(mode, k, qb, ip, tt, ww)     # host-buffer path (re-stages slot 0)
        assert np.array_equal(oi.cpu().numpy(), ids) and np.array_equal(osc.cpu().numpy(), scores)
        assert np.array_equal(oc.cpu().numpy()[:B], counts) and oc.cpu().numpy()[B] == 0
    bad = t_dev.clone()
    bad[1] = bad[0]                                                    # duplicate index inside query 0
    with pytest.raises(B200RagError):
        sh.stage_device("sparse", k, B, None, ip, bad, w_dev)
    x[0, 3] = float("nan")
    with pytest.raises(B200RagError):
        sh.normalize_bf16_device(x, B, bits)
    sh.close()
