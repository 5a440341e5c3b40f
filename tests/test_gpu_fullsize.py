"""Full-size checks (BASELINE.json sizes: 10M x 1024 + Zipf postings on one B200).

  * FULL-CORPUS ORACLE: the shard's stored rows are streamed back to the host and ALL 10M rows / documents are scored
    by the C oracle for every checked query (tests/fullscale.py); the engine's complete per-leg top-L id lists, the
    fused hybrid ids and every score must equal the oracle's bit for bit.  Covered: dense / sparse / hybrid top-10,
    hybrid top-100 (L = 200 per leg), a 256-token (HyDE-length) sparse query, and BASELINE config 4's shape -- 1 000
    Zipf-skewed collections, one 64-query hybrid batch with a per-query collection bitmask (largest, a mid-sized and a
    small tenant checked against the oracle);
  * the stored rows are the generators' rows (a 64k-row window regenerated on the host);
  * planted dense queries retrieve their planted row first (the generator's known answer);
  * a 2-shard split of the same corpus, fused, equals the single shard bit for bit (sharding invariance);
  * the tcgen05 path, the SIMT path and the exhaustive exact path return identical results;
  * every batched kernel returns what the single-query kernels return.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROWS = int(os.environ.get("B200RAG_FULLSIZE_ROWS", 10_000_000))
N_COLL = 1000


@pytest.fixture(scope="module")
def big(gpu):
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from probe import build_shard
    dev = torch.device("cuda", gpu)
    torch.cuda.set_device(dev)
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 130e9 and ROWS >= 10_000_000:
        pytest.skip("not enough free HBM for the full-size fixture")
    one = build_shard(ROWS, 1024, True, dev)
    half = ROWS // 2 // 8192 * 8192
    a = build_shard(half, 1024, True, dev, n_total=ROWS, row0=0)
    b = build_shard(ROWS - half, 1024, True, dev, n_total=ROWS, row0=half)
    yield one, a, b, dev
    for s in (one, a, b):
        s.close()


def _eq(got_ids, got_sc, got_cnt, exp_ids, exp_sc, ctx):
    assert int(got_cnt) == len(exp_ids), f"{ctx}: count {int(got_cnt)} != {len(exp_ids)}"
    assert np.array_equal(got_ids[:len(exp_ids)], exp_ids), f"{ctx}: ids differ\n got {got_ids[:len(exp_ids)]}\n exp {exp_ids}"
    assert np.array_equal(np.asarray(got_sc[:len(exp_ids)], np.float64), np.asarray(exp_sc, np.float64)), f"{ctx}: scores differ"


def test_fullsize_full_corpus_oracle(big):
    """Every checked query: the oracle scores ALL rows; complete leg lists, fused ids and scores compared at tolerance 0."""
    import torch
    from b200rag import normalize_bf16, synth
    from fullscale import LegJob, stream_oracle_legs
    from oracle import fast
    one, _, _, dev = big
    thr = synth.zipf_thresholds(synth.VOCAB)

    # ---- queries
    nq = 3
    qf = synth.dense_queries_f32(2000, 0, nq, ROWS, 1024, corpus_seed=1234)
    ip, tt, ww = synth.sparse_queries(2000, 0, nq)
    qb = normalize_bf16(qf)
    lf = synth.dense_queries_f32(2000, 500, 1, ROWS, 1024, corpus_seed=1234)          # HyDE-length query: 256 tokens
    lip, ltt, lww = synth.sparse_queries(2000, 500, 1, n_tokens=256)
    lb = normalize_bf16(lf)
    assert lip[1] > 100, "the long query should have well over 100 distinct terms"
    # config 4: a batch of 64 queries, each against its own collection (Zipf over 1 000 tenants, like the rows)
    B4 = 64
    cf = synth.dense_queries_f32(2000, 1000, B4, ROWS, 1024, corpus_seed=1234)
    cip, ctt, cww = synth.sparse_queries(2000, 1000, B4)
    cb = normalize_bf16(cf)
    colls = synth.row_collections(99, 0, B4, N_COLL)
    cthr = synth.zipf_thresholds(N_COLL)
    cthr_dev = torch.from_numpy(cthr.view(np.int64)).to(dev)
    words = torch.empty((ROWS + 31) // 32, dtype=torch.int32, device=dev)
    mask_of = {}
    for c in sorted(set(int(x) for x in colls)):
        one.synth_collection_mask(1234, 0, ROWS, cthr_dev, N_COLL, c, words)
        one.mask_set(100 + c, words, ROWS)
        mask_of[c] = 100 + c
    del words
    order = np.argsort(colls, kind="stable")
    probe4 = sorted({int(order[0]), int(order[len(order) // 2]), int(order[-1])})   # largest, a mid-sized, a small tenant

    def elig_of(c):
        return lambda lo, hi: synth.row_collections(1234, lo, hi - lo, N_COLL, cthr) == c

    jobs = [LegJob(qb[i], tt[ip[i]:ip[i + 1]], ww[ip[i]:ip[i + 1]], L=20) for i in range(2)]
    jobs.append(LegJob(qb[2], tt[ip[2]:ip[3]], ww[ip[2]:ip[3]], L=200))                 # top-100 hybrid
    jobs.append(LegJob(lb[0], ltt, lww, L=20))                                           # 256-token query
    for i in probe4:
        jobs.append(LegJob(cb[i], ctt[cip[i]:cip[i + 1]], cww[cip[i]:cip[i + 1]], L=20, eligible=elig_of(int(colls[i]))))

    # ---- the stored rows are the generators' rows (one 64k window), then stream everything through the oracle
    win = (ROWS // 3) // 65536 * 65536
    tables = synth.bm25_tables(ROWS)

    def check_rows(lo, bits, rip, rtt, rww):
        if lo <= win < lo + len(bits):
            o = win - lo
            m = min(65536, len(bits) - o)
            assert np.array_equal(bits[o:o + m], fast.synth_dense_bf16(1234, win, m, 1024))
            gi, gt, gw = fast.synth_sparse_csr(1234, win, m, thr, tables[0], tables[1], synth.VOCAB, 256,
                                                synth.TERM_PERM_MUL % synth.VOCAB)
            assert np.array_equal(rip[o:o + m + 1] - rip[o], gi)
            assert np.array_equal(rtt[rip[o]:rip[o + m]], gt) and np.array_equal(rww[rip[o]:rip[o + m]], gw)

    stream_oracle_legs(one, jobs, check_rows=check_rows)

    # ---- plain top-10: each leg's COMPLETE list (depth 20 and 10), and the fusion
    tgt = synth.query_target_rows(2000, np.arange(nq), ROWS)
    for i in range(2):
        j = jobs[i]
        sl = (ip[i:i + 2] - ip[i], tt[ip[i]:ip[i + 1]], ww[ip[i]:ip[i + 1]])
        for k in (10, 20):
            ids, sc, cnt = one.search("dense", k, qb[i:i + 1])
            _eq(ids[0], sc[0], cnt[0], j.dense[0][:k], j.dense[1][:k], f"dense top-{k} q{i}")
            ids, sc, cnt = one.search("sparse", k, qb[i:i + 1], *sl)
            _eq(ids[0], sc[0], cnt[0], j.sparse[0][:k], j.sparse[1][:k], f"sparse top-{k} q{i}")
        assert j.dense[0][0] == tgt[i], "the planted row is the oracle's best dense hit"
        ids, sc, cnt = one.search("hybrid", 10, qb[i:i + 1], *sl)
        ei, es = j.hybrid(10)
        _eq(ids[0], sc[0], cnt[0], ei, es, f"hybrid top-10 q{i}")
        # the same through the tcgen05 path and through the exhaustive exact path
        one.set_dense_path(2)
        g = one.search("hybrid", 10, qb[i:i + 1], *sl)
        one.set_dense_path(0)
        _eq(g[0][0], g[1][0], g[2][0], ei, es, f"hybrid top-10 q{i} (tcgen05)")
    one.set_exhaustive(True)
    g = one.search("hybrid", 10, qb[0:1], ip[0:2], tt[ip[0]:ip[1]], ww[ip[0]:ip[1]])
    assert one.stats()["exhaustive"] == 1
    one.set_exhaustive(False)
    _eq(g[0][0], g[1][0], g[2][0], *jobs[0].hybrid(10), "hybrid top-10 q0 (exhaustive)")

    # ---- top-100 hybrid (legs of 200)
    j = jobs[2]
    sl = (ip[2:4] - ip[2], tt[ip[2]:ip[3]], ww[ip[2]:ip[3]])
    ids, sc, cnt = one.search("dense", 100, qb[2:3])
    _eq(ids[0], sc[0], cnt[0], j.dense[0][:100], j.dense[1][:100], "dense top-100")
    ids, sc, cnt = one.search("sparse", 100, qb[2:3], *sl)
    _eq(ids[0], sc[0], cnt[0], j.sparse[0][:100], j.sparse[1][:100], "sparse top-100")
    ids, sc, cnt = one.search("hybrid", 100, qb[2:3], *sl)
    _eq(ids[0], sc[0], cnt[0], *j.hybrid(100), "hybrid top-100")

    # ---- 256-token query
    j = jobs[3]
    ids, sc, cnt = one.search("sparse", 20, lb, lip, ltt, lww)
    _eq(ids[0], sc[0], cnt[0], j.sparse[0], j.sparse[1], "sparse top-20, 256-token query")
    ids, sc, cnt = one.search("hybrid", 10, lb, lip, ltt, lww)
    _eq(ids[0], sc[0], cnt[0], *j.hybrid(10), "hybrid top-10, 256-token query")

    # ---- config 4: ONE batch of 64 masked hybrid queries
    mids = np.asarray([mask_of[int(c)] for c in colls], dtype=np.int32)
    ids, sc, cnt = one.search("hybrid", 10, cb, cip, ctt, cww, mask_ids=mids)
    assert one.stats()["dense_path"] == 2
    for n_, i in enumerate(probe4):
        _eq(ids[i], sc[i], cnt[i], *jobs[4 + n_].hybrid(10), f"config-4 batch, query {i} (collection {int(colls[i])})")
    # every hit of every query of the batch belongs to the query's collection
    rc_all = synth.row_collections(1234, 0, ROWS, N_COLL, cthr)
    for i in range(B4):
        assert (rc_all[ids[i, :cnt[i]]] == colls[i]).all(), f"config-4 query {i}: a hit outside its collection"
    for m in mask_of.values():
        one.mask_drop(m)


def test_fullsize_properties(big):
    import torch
    from b200rag import Shard, normalize_bf16, synth
    one, a, b, dev = big
    assert one.count == ROWS and a.count + b.count == ROWS
    nq, k = 4, 10
    qf = synth.dense_queries_f32(2000, 0, nq, ROWS, 1024, corpus_seed=1234)
    ip, tt, ww = synth.sparse_queries(2000, 0, nq)
    qb = normalize_bf16(qf)
    tgt = synth.query_target_rows(2000, np.arange(nq), ROWS)

    ids, sc, cnt = one.search("dense", k, qb)
    for i in range(nq):
        assert cnt[i] == k and ids[i, 0] == tgt[i]
        assert all(sc[i, j] > sc[i, j + 1] or (sc[i, j] == sc[i, j + 1] and ids[i, j] < ids[i, j + 1])
                   for j in range(k - 1))
    ids_h, sc_h, cnt_h = one.search("hybrid", k, qb, ip, tt, ww)

    # ---- tcgen05 path == SIMT path
    one.set_dense_path(2)
    ids_g, sc_g, _ = one.search("hybrid", k, qb, ip, tt, ww)
    one.set_dense_path(0)
    assert np.array_equal(ids_g, ids_h) and np.array_equal(sc_g, sc_h)

    # ---- sharding invariance: legs per shard -> concatenate -> fuse == single shard
    q, keep = a.make_query("hybrid", k, qb, ip, tt, ww)
    nlegs, L = Shard.legs_len(q)
    n = nlegs * nq * L + 1
    gathered = torch.zeros((2, n, 2), dtype=torch.int64, device=dev)
    for r, sh in enumerate((a, b)):
        sh.stage(q, keep)
        sh.legs(gathered[r], gathered[r, -1])
        sh.sync()
    out = torch.zeros(2 * nq * k + (nq + 2) // 2 + 1, dtype=torch.int64, device=dev)
    a.fuse(gathered, 2, out[:nq * k], out[nq * k:2 * nq * k], out[2 * nq * k:], has_trailer=True)
    a.sync()
    h = out.cpu().numpy()
    assert np.array_equal(h[:nq * k].reshape(nq, k), ids_h)
    assert np.array_equal(h[nq * k:2 * nq * k].view(np.float64).reshape(nq, k), sc_h)
    assert h[2 * nq * k:].view(np.int32)[nq] == 0

    # ---- the same two halves as a shard GROUP (one process, worker threads): b200rag_group_search == single shard
    from b200rag import ShardGroup
    grp = ShardGroup([a, b])
    gi, gs, gc = grp.search("hybrid", k, qb, ip, tt, ww)
    assert np.array_equal(gi, ids_h) and np.array_equal(gs, sc_h) and np.array_equal(gc, cnt_h)
    gi, gs, gc = grp.search("dense", 100, qb)
    di, ds, dc = one.search("dense", 100, qb)
    assert np.array_equal(gi, di) and np.array_equal(gs, ds) and np.array_equal(gc, dc)
    grp.close()


def test_fullsize_batched_paths_agree_with_single_query_paths(big):
    """At 10M rows every batched kernel must return, bit for bit, what the single-query kernels return:
    tcgen05 on CTA pairs (160 queries, one pass of 256) vs the SIMT scan; the sample+filter epilogue (top-100) vs the
    SIMT scan; the multi-block sparse scan with 24 queries in flight vs one query at a time; hybrid batches likewise."""
    from b200rag import normalize_bf16, synth
    one, _, _, _ = big
    nq = 160
    qf = synth.dense_queries_f32(2000, 100, nq, ROWS, 1024, corpus_seed=1234)
    ip, tt, ww = synth.sparse_queries(2000, 100, nq)
    qb = normalize_bf16(qf)

    def single(mode, k, idx):
        out = []
        for i in idx:
            r = one.search(mode, k, qb[i:i + 1], ip[i:i + 2] - ip[i], tt[ip[i]:ip[i + 1]], ww[ip[i]:ip[i + 1]])
            assert one.stats()["dense_path"] in (0, 1)
            out.append((r[0][0], r[1][0], int(r[2][0])))
        return out

    probe = [0, 1, 77, 127, 128, 159]
    # CTA pairs, register lists
    ids, sc, cnt = one.search("dense", 10, qb)
    assert one.stats()["dense_path"] == 2 and one.stats()["dense_passes"] == 1
    for i, (ei, es, ec) in zip(probe, single("dense", 10, probe)):
        assert cnt[i] == ec and np.array_equal(ids[i], ei) and np.array_equal(sc[i], es)
    # sample + filter epilogue (top-100), also on pairs
    ids, sc, cnt = one.search("dense", 100, qb)
    assert one.stats()["dense_passes"] == 1 and one.stats()["retries"] == 0
    for i, (ei, es, ec) in zip(probe[:3], single("dense", 100, probe[:3])):
        assert cnt[i] == ec and np.array_equal(ids[i], ei) and np.array_equal(sc[i], es)
    # sparse and hybrid batches
    sl = slice(0, 24)
    sip = ip[:25] - ip[0]
    ids, sc, cnt = one.search("sparse", 10, qb[sl], sip, tt[:ip[24]], ww[:ip[24]])
    for i, (ei, es, ec) in zip([0, 5, 23], single("sparse", 10, [0, 5, 23])):
        assert cnt[i] == ec and np.array_equal(ids[i, :ec], ei[:ec]) and np.array_equal(sc[i, :ec], es[:ec])
    ids, sc, cnt = one.search("hybrid", 10, qb[sl], sip, tt[:ip[24]], ww[:ip[24]])
    for i, (ei, es, ec) in zip([0, 5, 23], single("hybrid", 10, [0, 5, 23])):
        assert cnt[i] == ec and np.array_equal(ids[i, :ec], ei[:ec]) and np.array_equal(sc[i, :ec], es[:ec])
