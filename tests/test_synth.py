"""Synthetic generators: determinism, shard independence, distribution sanity."""
import numpy as np

from b200rag import synth


def test_rows_are_keyed_by_global_row_id():
    a = synth.dense_rows_bf16(1234, 0, 64, 256)
    b = synth.dense_rows_bf16(1234, 32, 32, 256)
    assert np.array_equal(a[32:], b)
    assert not np.array_equal(a, synth.dense_rows_bf16(1235, 0, 64, 256))
    f = synth.bf16_bits_to_f32(a).astype(np.float64)
    assert np.allclose(np.sqrt((f * f).sum(1)), 1.0, atol=5e-3)


def test_bf16_rounding_is_rne():
    y = np.array([1.0, 1.00390625, 1.01171875, -2.5, 3.0e-5], np.float32)   # 1+2^-8 is a tie -> even
    bits = synth.f32_to_bf16_bits(y)
    back = synth.bf16_bits_to_f32(bits)
    assert back[0] == 1.0 and back[1] == 1.0 and back[2] == np.float32(1.015625) and back[3] == -2.5


def test_planted_queries_hit_their_row():
    n = 500
    bits = synth.dense_rows_bf16(7, 0, n, 256)
    q = synth.dense_queries_f32(9, 0, 20, n, 256, corpus_seed=7)
    tgt = synth.query_target_rows(9, np.arange(20), n)
    sc = synth.bf16_bits_to_f32(bits) @ q.T
    for i in range(20):
        if i % 10 != 9:
            assert int(np.argmax(sc[:, i])) == int(tgt[i])


def test_zipf_sparse_shapes():
    V = synth.VOCAB
    thr = synth.zipf_thresholds(V)
    ip, tt, ww = synth.sparse_docs_csr(3, 100, 300, 1_000_000, thresholds=thr)
    nnz = np.diff(ip)
    assert 170 < nnz.mean() < 215                       # expected 193.9 distinct terms / 256-token chunk
    for d in range(0, 300, 37):
        t = tt[ip[d]:ip[d + 1]]
        assert (np.diff(t.astype(np.int64)) > 0).all() and t.max() < V
    assert (ww > 0).all()
    ip2, tt2, ww2 = synth.sparse_docs_csr(3, 200, 50, 1_000_000, thresholds=thr)
    assert np.array_equal(tt[ip[100]:ip[150]], tt2) and np.array_equal(ww[ip[100]:ip[150]], ww2)
    qi, qt, qw = synth.sparse_queries(5, 0, 50, 12, V, thr)
    assert (np.diff(qi) <= 12).all() and qw.sum() == 50 * 12


def test_collections_and_mask_packing():
    c = synth.row_collections(1, 0, 5000, 1000)
    assert c.min() >= 0 and c.max() < 1000 and (c == 0).mean() > 0.08   # Zipf head tenant
    bits = np.zeros(70, bool)
    bits[[0, 31, 32, 69]] = True
    w = synth.pack_mask(bits)
    assert w.dtype == np.uint32 and len(w) == 3
    assert w[0] == (1 | (1 << 31)) and w[1] == 1 and w[2] == (1 << 5)
