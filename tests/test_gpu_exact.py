"""GPU tests of the exactness contract and of the row-level operations added for it:

  * unresolved ambiguity is LOUD: a corpus of exact duplicates around the cut can never clear the slack guard; with the
    exhaustive fallback disabled `b200rag_search` fails with B200RAG_ERR_INEXACT, with it enabled (the default) the
    result equals the oracle's (duplicates ordered by smaller row id, rule R5);
  * the exhaustive exact path (`set_exhaustive`) is an independent implementation of the legs -- no scan kernel, no
    candidate cut -- and must agree with the oracle and with the scan path in every mode, with masks and thresholds;
  * explicit global row ids (`add(..., ids=)`): results carry them, ties break on them, non-increasing ids are refused;
  * `add_f32` (GPU-side normalise + pack) stores the same bits as the host routine;
  * `compact` drops rows physically: the survivors keep ids and order, searches equal the oracle on the survivors,
    rows can be added afterwards; `read_sparse` / `read_row_ids` return what was stored;
  * shard files (format v2) keep explicit ids.
"""
import numpy as np
import pytest

from helpers import Corpus, assert_result_equal, oracle_search

pytestmark = pytest.mark.gpu


def _dup_corpus(n=6000, dup_from=100, dup_count=2000, dim=1024, vocab=50_021):
    """Rows [dup_from, dup_from + dup_count) are exact copies of row `dup_from` (dense bits AND sparse vector)."""
    c = Corpus(n, dim=dim, vocab=vocab)
    ip, tt, ww = c.indptr, c.terms, c.w
    s0, e0 = ip[dup_from], ip[dup_from + 1]
    lens = np.diff(ip).copy()
    lens[dup_from:dup_from + dup_count] = e0 - s0
    new_ip = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    new_t = np.empty(new_ip[-1], np.uint32)
    new_w = np.empty(new_ip[-1], np.float32)
    for d in range(n):
        src = slice(s0, e0) if dup_from <= d < dup_from + dup_count else slice(ip[d], ip[d + 1])
        new_t[new_ip[d]:new_ip[d + 1]] = tt[src]
        new_w[new_ip[d]:new_ip[d + 1]] = ww[src]
    c.bits[dup_from:dup_from + dup_count] = c.bits[dup_from]
    c.indptr, c.terms, c.w = new_ip, new_t, new_w
    return c


def test_unresolved_ambiguity_is_loud_or_exhaustive(gpu):
    from b200rag import B200RagError, Shard, _ffi, normalize_bf16, synth
    c = _dup_corpus()
    sh = Shard(dim=c.dim, vocab=c.vocab, device=gpu, docs_per_block=2048)
    sh.add(c.bits, c.indptr, c.terms, c.w)
    # a query whose best rows are the 2000 duplicates: their own vector, their own terms
    qb = c.bits[100:101].copy()
    sl = slice(c.indptr[100], c.indptr[101])
    q_ip = np.asarray([0, sl.stop - sl.start], dtype=np.int64)
    q_t, q_w = c.terms[sl].copy(), np.ones(sl.stop - sl.start, np.float32)
    for mode in ("dense", "sparse", "hybrid"):
        sh.set_exact_fallback(False)
        with pytest.raises(B200RagError) as ei:
            sh.search(mode, 10, qb, q_ip, q_t, q_w)
        assert ei.value.code == _ffi.ERR_INEXACT, "a result whose guard never cleared must not be returned as OK"
        sh.set_exact_fallback(True)
        ids, sc, cnt = sh.search(mode, 10, qb, q_ip, q_t, q_w)
        st = sh.stats()
        assert st["exhaustive"] == 1 and st["retries"] >= 1
        e_i, e_s = oracle_search(c, mode, qb[0], q_t, q_w, None, 10)
        assert_result_equal(ids[0], sc[0], int(cnt[0]), e_i, e_s, ctx=f"{mode} over 2000 exact duplicates")
        assert list(ids[0]) == list(range(100, 110)), "exact ties resolve to the smaller row ids (R5)"
    # an ordinary query on the same shard needs neither retries nor the fallback
    qf, ip, tt, ww = c.queries(1, qid_start=50)
    ids, sc, cnt = sh.search("hybrid", 10, normalize_bf16(qf), ip, tt, ww)
    assert sh.stats()["exhaustive"] == 0
    sh.close()


@pytest.mark.parametrize("dim", [256, 1024])
def test_exhaustive_path_equals_oracle_and_scan_path(gpu, dim):
    from b200rag import Shard, normalize_bf16
    c = Corpus(9000, dim=dim, vocab=40_009)
    sh = Shard(dim=dim, vocab=c.vocab, device=gpu, docs_per_block=1024, row_base=5000)
    sh.add(c.bits, c.indptr, c.terms, c.w)
    rng = np.random.default_rng(3)
    masks = {0: rng.random(c.n) < 0.3, 1: np.zeros(c.n, bool), 2: rng.random(c.n) < 0.01}
    from b200rag.synth import pack_mask
    for m, bits in masks.items():
        sh.mask_set(m, pack_mask(bits), c.n)
    qf, ip, tt, ww = c.queries(5)
    qb = normalize_bf16(qf)
    mids = np.asarray([0, -1, 1, 2, 0], dtype=np.int32)
    for mode, k, thr in (("dense", 10, None), ("dense", 100, 0.05), ("sparse", 7, None), ("hybrid", 10, None),
                         ("hybrid", 100, None)):
        fast_r = sh.search(mode, k, qb, ip, tt, ww, mask_ids=mids, score_threshold=thr)
        sh.set_exhaustive(True)
        slow_r = sh.search(mode, k, qb, ip, tt, ww, mask_ids=mids, score_threshold=thr)
        assert sh.stats()["exhaustive"] == 1
        sh.set_exhaustive(False)
        for a, b in zip(fast_r, slow_r):
            assert np.array_equal(a, b), f"{mode} k={k}: scan path and exhaustive path differ"
        for b in range(5):
            elig = None if mids[b] < 0 else masks[int(mids[b])]
            e_i, e_s = oracle_search(c, mode, qb[b], tt[ip[b]:ip[b + 1]], ww[ip[b]:ip[b + 1]], elig, k, thr, row_base=5000)
            assert_result_equal(slow_r[0][b], slow_r[1][b], int(slow_r[2][b]), e_i, e_s, ctx=f"exhaustive {mode} q{b}")
    sh.close()


def test_explicit_row_ids_add_f32_and_read_back(gpu):
    from b200rag import B200RagError, Shard, normalize_bf16, synth
    c = Corpus(3000, dim=512, vocab=30_011)
    f32 = synth.bf16_bits_to_f32(c.bits) * np.linspace(0.5, 3.0, c.n, dtype=np.float32)[:, None]   # un-normalised input
    bits = normalize_bf16(f32)
    ids = np.cumsum(np.random.default_rng(0).integers(1, 5, size=c.n)).astype(np.int64) + 1000     # increasing, gappy
    a = Shard(dim=512, vocab=c.vocab, device=gpu, docs_per_block=1024)
    b = Shard(dim=512, vocab=c.vocab, device=gpu, docs_per_block=1024)
    for s in range(0, c.n, 700):
        e = min(c.n, s + 700)
        sl = (c.indptr[s:e + 1] - c.indptr[s], c.terms[c.indptr[s]:c.indptr[e]], c.w[c.indptr[s]:c.indptr[e]])
        a.add(bits[s:e], *sl, ids=ids[s:e])
        b.add_f32(f32[s:e], *sl, ids=ids[s:e])
    assert np.array_equal(a.read_dense(0, c.n), bits) and np.array_equal(b.read_dense(0, c.n), bits), \
        "GPU-side normalise + pack stores the host routine's bits"
    for sh in (a, b):
        assert np.array_equal(sh.read_row_ids(0, c.n), ids)
        rip, rt, rw = sh.read_sparse(500, 1200)
        assert np.array_equal(rip, c.indptr[500:1701] - c.indptr[500])
        assert np.array_equal(rt, c.terms[c.indptr[500]:c.indptr[1700]]) and np.array_equal(rw, c.w[c.indptr[500]:c.indptr[1700]])
    with pytest.raises(B200RagError):
        a.add(bits[:2], ids=np.asarray([ids[-1], ids[-1] + 1]))          # not above the ids stored so far
    with pytest.raises(B200RagError):
        a.add(bits[:2], ids=np.asarray([ids[-1] + 5, ids[-1] + 5]))      # not strictly increasing
    c.bits = bits
    qf, ip, tt, ww = c.queries(3)
    qb = normalize_bf16(qf)
    for mode in ("dense", "sparse", "hybrid"):
        ra = a.search(mode, 10, qb, ip, tt, ww)
        rb = b.search(mode, 10, qb, ip, tt, ww)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
        for q in range(3):
            e_i, e_s = oracle_search(c, mode, qb[q], tt[ip[q]:ip[q + 1]], ww[ip[q]:ip[q + 1]], None, 10)
            assert_result_equal(ra[0][q], ra[1][q], int(ra[2][q]), ids[e_i], e_s, ctx=f"explicit ids {mode} q{q}")
    a.close()
    b.close()


def test_compact_drops_rows_physically(gpu, tmp_path):
    from b200rag import B200RagError, Shard, normalize_bf16
    from b200rag.synth import pack_mask
    c = Corpus(12_000, dim=256, vocab=30_011)
    sh = Shard(dim=256, vocab=c.vocab, device=gpu, docs_per_block=1024, row_base=100)
    sh.add(c.bits, c.indptr, c.terms, c.w)
    qf, ip, tt, ww = c.queries(4)
    qb = normalize_bf16(qf)
    sh.search("hybrid", 10, qb, ip, tt, ww)                     # builds the inverted index before the compaction
    rng = np.random.default_rng(5)
    keep = rng.random(c.n) < 0.6
    keep[:1500] = False                                          # a whole leading block disappears
    with pytest.raises(B200RagError):
        sh.compact(pack_mask(keep[:-1]), c.n - 1)
    sh.mask_set(3, pack_mask(np.ones(c.n, bool)), c.n)
    sh.compact(pack_mask(keep), c.n)
    n2 = int(keep.sum())
    assert sh.count == n2 and sh.postings == int(np.diff(c.indptr)[keep].sum())
    kept = np.flatnonzero(keep)
    assert np.array_equal(sh.read_row_ids(0, n2), kept + 100), "survivors keep their global ids and their order"
    assert np.array_equal(sh.read_dense(0, n2), c.bits[keep])
    with pytest.raises(B200RagError):
        sh.search("dense", 5, qb, mask_ids=np.asarray([3, 3, 3, 3], np.int32))   # masks were dropped with the rows
    for mode, k in (("dense", 10), ("sparse", 10), ("hybrid", 10), ("hybrid", 100)):
        ids, sc, cnt = sh.search(mode, k, qb, ip, tt, ww)
        for q in range(4):
            e_i, e_s = oracle_search(c, mode, qb[q], tt[ip[q]:ip[q + 1]], ww[ip[q]:ip[q + 1]], keep, k, row_base=100)
            assert_result_equal(ids[q], sc[q], int(cnt[q]), e_i, e_s, ctx=f"after compact {mode} k={k} q{q}")
    # rows added after a compaction continue the id space; implicit ids must stay above the stored ones
    with pytest.raises(B200RagError):
        sh.add(c.bits[:4])                                       # row_base + local row would reuse dropped ids
    extra = Corpus(500, dim=256, vocab=30_011, seed=77)
    new_ids = np.arange(100 + c.n, 100 + c.n + 500, dtype=np.int64)
    sh.add(extra.bits, extra.indptr, extra.terms, extra.w, ids=new_ids)
    c2 = Corpus(1, dim=256, vocab=30_011)                       # container for the concatenated oracle corpus
    c2.n = n2 + 500
    c2.bits = np.concatenate([c.bits[keep], extra.bits])
    lens = np.concatenate([np.diff(c.indptr)[keep], np.diff(extra.indptr)])
    c2.indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pos = np.repeat(keep, np.diff(c.indptr))
    c2.terms = np.concatenate([c.terms[pos], extra.terms])
    c2.w = np.concatenate([c.w[pos], extra.w])
    gid = np.concatenate([kept + 100, new_ids])
    ids, sc, cnt = sh.search("hybrid", 10, qb, ip, tt, ww)
    for q in range(4):
        e_i, e_s = oracle_search(c2, "hybrid", qb[q], tt[ip[q]:ip[q + 1]], ww[ip[q]:ip[q + 1]], None, 10)
        assert_result_equal(ids[q], sc[q], int(cnt[q]), gid[e_i], e_s, ctx=f"compact + add q{q}")
    # shard file v2 keeps the explicit ids
    path = str(tmp_path / "shard.bin")
    sh.save(path)
    back = Shard(dim=256, vocab=c.vocab, device=gpu, docs_per_block=1024, row_base=0)
    back.load(path)
    assert np.array_equal(back.read_row_ids(0, back.count), gid)
    r2 = back.search("hybrid", 10, qb, ip, tt, ww)
    assert np.array_equal(r2[0], ids) and np.array_equal(r2[1], sc)
    # compacting everything away leaves an empty, usable shard
    back.compact(pack_mask(np.zeros(back.count, bool)), back.count)
    assert back.count == 0 and back.search("hybrid", 5, qb, ip, tt, ww)[2].sum() == 0
    back.add(extra.bits, extra.indptr, extra.terms, extra.w, ids=new_ids + 10_000)
    assert back.search("dense", 5, qb)[2].min() == 5
    sh.close()
    back.close()
