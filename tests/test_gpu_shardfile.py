"""GPU side of b200rag/shardfile.py (CPU side: tests/test_shardfile.py): what b200rag_save writes IS the documented flat
layout, and what the Python writer produces loads into the engine."""
import numpy as np
import pytest

from helpers import Corpus

pytestmark = pytest.mark.gpu


def test_shard_file_is_the_documented_layout(gpu, tmp_path):
    """b200rag_save writes exactly the flat, mmap-able layout b200rag/shardfile.py documents: the engine's file is byte
    for byte what the Python writer produces for the same rows, numpy.memmap reads it back without a device, and the
    engine loads a Python-written file (an ingest job on a CPU-only machine) and answers like the shard that saved."""
    from b200rag import Shard, normalize_bf16, shardfile
    c = Corpus(5_000, dim=1024, vocab=40_009)
    ids = np.arange(c.n, dtype=np.int64) * 2 + 5
    sh = Shard(dim=1024, vocab=c.vocab, device=gpu, docs_per_block=2048)
    sh.add(c.bits, c.indptr, c.terms, c.w, ids=ids)
    p_engine, p_python = str(tmp_path / "engine.bin"), str(tmp_path / "python.bin")
    sh.save(p_engine)
    f = shardfile.open_mmap(p_engine)
    assert (f.version, f.dim, f.vocab, f.n_rows, f.nnz) == (2, 1024, c.vocab, c.n, len(c.terms))
    assert np.array_equal(f.dense, c.bits) and np.array_equal(f.indptr, c.indptr) and np.array_equal(f.terms, c.terms)
    assert np.array_equal(f.weights, c.w) and np.array_equal(f.row_ids, ids)
    shardfile.write(p_python, c.bits, c.indptr, c.terms, c.w, ids, vocab=c.vocab)
    assert open(p_engine, "rb").read() == open(p_python, "rb").read()
    back = Shard(dim=1024, vocab=c.vocab, device=gpu, docs_per_block=2048)
    back.load(p_python)
    qf, ip, tt, ww = c.queries(3)
    qb = normalize_bf16(qf)
    for mode in ("dense", "sparse", "hybrid"):
        a = sh.search(mode, 10, qb, ip, tt, ww)
        b = back.search(mode, 10, qb, ip, tt, ww)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), mode
    assert np.array_equal(back.read_row_ids(0, c.n), ids)
    sh.close()
    back.close()
