"""GPU tests of the opt-in 8-bit candidate scan (csrc/dense_q8.cu): the scan reads an int8 copy of the
rows and ranks them by a rigorous upper bound of their exact score; the candidates are re-scored from the bf16 rows, so
ids AND scores must stay bit-identical to the oracle (and to the uncompressed path), whatever the data:
ordinary rows, rows with one dominant component (a large quantisation step), exact duplicates, masks, thresholds,
two queries per pass, rows appended / compacted / reloaded after compression was switched on."""
import numpy as np
import pytest

from helpers import Corpus, assert_result_equal, oracle_search

pytestmark = pytest.mark.gpu


def _check(sh, c, mode, qb, ip, tt, ww, k, elig=None, mids=None, thr=None, row_base=0, expect_q8=True, ctx=""):
    r = sh.search(mode, k, qb, ip, tt, ww, mask_ids=mids, score_threshold=thr)
    st = sh.stats()
    if expect_q8 and mode != "sparse":
        assert st["dense_path"] == 3 or st["retries"] > 0, f"{ctx}: the 8-bit scan was not taken"
    for b in range(len(qb)):
        e = None if elig is None else elig[b]
        e_i, e_s = oracle_search(c, mode, qb[b], tt[ip[b]:ip[b + 1]], ww[ip[b]:ip[b + 1]], e, k, thr, row_base=row_base)
        assert_result_equal(r[0][b], r[1][b], int(r[2][b]), e_i, e_s, ctx=f"{ctx} {mode} k={k} q{b}")
    return st


@pytest.mark.parametrize("dim", [512, 1024])
def test_q8_scan_is_exact(gpu, dim):
    from b200rag import Shard, normalize_bf16
    from b200rag.synth import pack_mask
    c = Corpus(60_000, dim=dim, vocab=60_013)
    # adversarial rows: one dominant component (scale = 1/127: a wide error band), near-duplicates of a planted target
    rng = np.random.default_rng(11)
    f = np.zeros((300, dim), np.float32)
    f[np.arange(300), rng.integers(0, dim, 300)] = 1.0
    f += rng.standard_normal((300, dim)).astype(np.float32) * 0.02
    c.bits[1000:1300] = normalize_bf16(f)
    c.bits[5000:5040] = c.bits[4999]                         # 40 exact duplicates
    sh = Shard(dim=dim, vocab=c.vocab, device=gpu, docs_per_block=2048, row_base=7)
    sh.add(c.bits[:30_000], c.indptr[:30_001], c.terms[:c.indptr[30_000]], c.w[:c.indptr[30_000]])
    sh.set_compression(True)                                 # quantises the rows stored so far ...
    lo = 30_000
    sh.add(c.bits[lo:], c.indptr[lo:] - c.indptr[lo], c.terms[c.indptr[lo]:], c.w[c.indptr[lo]:])   # ... and every later add
    qf, ip, tt, ww = c.queries(6)
    qb = normalize_bf16(qf)
    qb[4] = c.bits[4999]                                     # the duplicated row itself
    qb[5] = c.bits[1100]                                     # one of the dominant-component rows
    one = [slice(b, b + 1) for b in range(6)]
    for b in range(6):                                       # one query per pass
        sl = (ip[b:b + 2] - ip[b], tt[ip[b]:ip[b + 1]], ww[ip[b]:ip[b + 1]])
        for mode, k in (("dense", 10), ("hybrid", 10), ("dense", 100)):
            _check(sh, c, mode, qb[one[b]], *sl, k, row_base=7, ctx=f"q8 B=1 q{b}")
    # two queries per pass, masks, threshold
    masks = {0: rng.random(c.n) < 0.3, 1: rng.random(c.n) < 0.004}
    for m, bits in masks.items():
        sh.mask_set(m, pack_mask(bits), c.n)
    sl2 = (ip[0:3] - ip[0], tt[ip[0]:ip[2]], ww[ip[0]:ip[2]])
    _check(sh, c, "hybrid", qb[0:2], *sl2, 10, elig=[masks[0], masks[1]], mids=np.asarray([0, 1], np.int32), row_base=7, ctx="q8 B=2 masks")
    _check(sh, c, "dense", qb[0:2], *sl2, 20, thr=0.08, row_base=7, ctx="q8 B=2 threshold")
    # same answers with compression off, and the stats say which scan ran
    a = sh.search("hybrid", 10, qb[0:1], ip[0:2], tt[:ip[1]], ww[:ip[1]])
    assert sh.stats()["dense_path"] == 3 and sh.stats()["dense_bytes"] == c.n * (dim + 16)
    sh.set_compression(False)
    b_ = sh.search("hybrid", 10, qb[0:1], ip[0:2], tt[:ip[1]], ww[:ip[1]])
    assert sh.stats()["dense_path"] == 1 and sh.stats()["dense_bytes"] == c.n * dim * 2
    assert all(np.array_equal(x, y) for x, y in zip(a, b_))
    sh.close()


def test_q8_follows_compaction_and_reload(gpu, tmp_path):
    from b200rag import Shard, normalize_bf16
    from b200rag.synth import pack_mask
    c = Corpus(20_000, dim=1024, vocab=40_009)
    sh = Shard(dim=1024, vocab=c.vocab, device=gpu, docs_per_block=2048)
    sh.set_compression(True)                                 # on an empty shard
    sh.add(c.bits, c.indptr, c.terms, c.w)
    qf, ip, tt, ww = c.queries(3)
    qb = normalize_bf16(qf)
    keep = np.random.default_rng(2).random(c.n) < 0.7
    sh.compact(pack_mask(keep), c.n)
    for b in range(3):
        sl = (ip[b:b + 2] - ip[b], tt[ip[b]:ip[b + 1]], ww[ip[b]:ip[b + 1]])
        _check(sh, c, "hybrid", qb[b:b + 1], *sl, 10, elig=[keep], ctx="q8 after compact")
    path = str(tmp_path / "s.bin")
    sh.save(path)
    back = Shard(dim=1024, vocab=c.vocab, device=gpu, docs_per_block=2048)
    back.set_compression(True)
    back.load(path)
    r1 = sh.search("dense", 10, qb[:1])
    r2 = back.search("dense", 10, qb[:1])
    assert back.stats()["dense_path"] == 3 and all(np.array_equal(x, y) for x, y in zip(r1, r2))
    sh.close()
    back.close()


def test_q8_guard_failure_retries_on_bf16(gpu):
    """3 000 near-copies of one row, closer to each other than the 8-bit error band is wide: the 8-bit scan cannot
    separate them (its weakest retained upper bound stays above the L-th exact score), says so, and the search is
    answered by the bf16 scan (or, if that is ambiguous too, exhaustively) -- never by the unproven candidate set."""
    from b200rag import Shard, normalize_bf16
    c = Corpus(30_000, dim=1024, vocab=40_009)
    rng = np.random.default_rng(5)
    target = rng.standard_normal(1024).astype(np.float32)
    target /= np.linalg.norm(target)
    near = target[None, :] + rng.standard_normal((3000, 1024)).astype(np.float32) * (0.05 / 32.0)
    c.bits[10_000:13_000] = normalize_bf16(near)
    sh = Shard(dim=1024, vocab=c.vocab, device=gpu, docs_per_block=2048)
    sh.add(c.bits, c.indptr, c.terms, c.w)
    sh.set_compression(True)
    qb = normalize_bf16(target[None, :])
    qf, ip, tt, ww = c.queries(1)
    for mode, k in (("dense", 10), ("hybrid", 10)):
        st = _check(sh, c, mode, qb, ip, tt, ww, k, ctx="q8 near-copies")
        assert st["retries"] > 0, "3 000 rows inside the error band fit no candidate list: the guard must have fired"
    # an ordinary query right after: the 8-bit scan again, first attempt
    qb2 = normalize_bf16(qf)
    st = _check(sh, c, "dense", qb2, ip, tt, ww, 10, ctx="q8 after a retry")
    assert st["dense_path"] == 3 and st["retries"] == 0
    sh.close()
