"""b200rag/shardfile.py on the CPU: the documented layout of the engine's shard file, written and mapped back without a
GPU (the engine-written file is compared with it in tests/test_gpu_shardfile.py::test_shard_file_is_the_documented_layout)."""
import os
import struct

import numpy as np
import pytest

from b200rag import shardfile
from helpers import Corpus


def test_write_then_mmap_round_trip(tmp_path):
    c = Corpus(700, dim=256, vocab=20_011)
    ids = np.arange(c.n, dtype=np.int64) * 3 + 11
    path = str(tmp_path / "s.bin")
    shardfile.write(path, c.bits, c.indptr, c.terms, c.w, ids, vocab=c.vocab)
    assert not os.path.exists(path + ".tmp")
    f = shardfile.open_mmap(path)
    assert (f.version, f.dim, f.vocab, f.n_rows, f.nnz) == (2, 256, c.vocab, c.n, len(c.terms))
    assert isinstance(f.dense, np.memmap) and f.dense.shape == (c.n, 256)
    assert np.array_equal(f.dense, c.bits) and np.array_equal(f.indptr, c.indptr)
    assert np.array_equal(f.terms, c.terms) and np.array_equal(f.weights, c.w) and np.array_equal(f.row_ids, ids)
    # byte-level: header fields at the documented offsets, sections back to back
    raw = open(path, "rb").read()
    assert raw[:8] == b"B200RAG1" and struct.unpack_from("<4i2q", raw, 8) == (2, 256, c.vocab, 0, c.n, len(c.terms))
    assert len(raw) == 40 + c.n * 256 * 2 + (c.n + 1) * 8 + len(c.terms) * 8 + c.n * 8
    assert raw[40:40 + 512] == c.bits[0].tobytes()
    assert raw[-8:] == struct.pack("<q", int(ids[-1]))


def test_empty_shard_dense_only_and_version_1(tmp_path):
    p = str(tmp_path / "e.bin")
    shardfile.write(p, np.zeros((0, 1024), np.uint16), None, None, None)
    f = shardfile.open_mmap(p)
    assert f.n_rows == 0 and f.nnz == 0 and f.indptr.tolist() == [0] and f.dense.shape == (0, 1024)
    c = Corpus(50, dim=256, vocab=20_011)
    shardfile.write(p, c.bits, None, None, None, row_base=500, vocab=c.vocab)      # dense-only rows, implicit ids
    f = shardfile.open_mmap(p)
    assert f.nnz == 0 and np.array_equal(f.indptr, np.zeros(51, np.int64)) and f.row_ids.tolist() == list(range(500, 550))
    # a version-1 file (round 1's writer): no id section
    with open(p, "wb") as fh:
        fh.write(shardfile.HEADER.pack(shardfile.MAGIC, 1, 256, c.vocab, 0, c.n, len(c.terms)))
        for a in (c.bits, c.indptr, c.terms, c.w):
            np.ascontiguousarray(a).tofile(fh)
    f = shardfile.open_mmap(p)
    assert f.version == 1 and f.row_ids is None and np.array_equal(f.terms, c.terms)


def test_refuses_what_the_engine_would_misread(tmp_path):
    c = Corpus(40, dim=256, vocab=20_011)
    p = str(tmp_path / "x.bin")
    with pytest.raises(ValueError):
        shardfile.write(p, c.bits, c.indptr[:-1], c.terms, c.w)
    with pytest.raises(ValueError):
        shardfile.write(p, c.bits, c.indptr, c.terms, c.w, np.zeros(c.n, np.int64))            # ids not increasing
    with pytest.raises(ValueError):
        shardfile.write(p, c.bits, c.indptr, c.terms, c.w, vocab=int(c.terms.max()))           # term outside the vocabulary
    t2 = c.terms.copy()
    a = int(c.indptr[3])
    t2[a], t2[a + 1] = t2[a + 1], t2[a]                                                         # a row's terms out of order
    with pytest.raises(ValueError):
        shardfile.write(p, c.bits, c.indptr, t2, c.w, vocab=c.vocab)
    shardfile.write(p, c.bits, c.indptr, c.terms, c.w, vocab=c.vocab)
    with open(p, "r+b") as fh:
        fh.truncate(os.path.getsize(p) - 16)
    with pytest.raises(ValueError):
        shardfile.open_mmap(p)
    with open(p, "wb") as fh:
        fh.write(b"not a shard")
    with pytest.raises(ValueError):
        shardfile.open_mmap(p)
