"""The on-disk shard format of ``b200rag_save`` / ``b200rag_load`` (csrc/engine.cu, ``ShardFileHeader``), readable and
writable WITHOUT a GPU.  SURVEY 8f rank 2 asks for an mmap-able format: the file is a 40-byte header followed by the
shard's arrays, raw and contiguous, so ``open_mmap`` maps them in place (cold start = page faults, no parsing), an
ingest job on a CPU-only machine can write a file the engine loads as is, and a checker can read a shard back without
a device.  The reference keeps its index in Qdrant's volume (docker-compose.yml:36-37) and has no file format of its own.

Layout (little endian, no padding between the sections):

    offset 0    char[8]  magic "B200RAG1"
           8    int32    version (2; version 1 files end after the weights: ids = row_base + local row)
           12   int32    dim
           16   int32    vocab
           20   int32    reserved (0)
           24   int64    n_rows
           32   int64    nnz
           40   uint16   dense[n_rows][dim]      bf16 bits of the unit rows
           ...  int64    indptr[n_rows + 1]      forward index, indptr[0] = 0, indptr[n_rows] = nnz
           ...  uint32   terms[nnz]              per row ascending, unique
           ...  float32  weights[nnz]
           ...  int64    row_ids[n_rows]         global id of every local row, strictly increasing (version >= 2)

The inverted index and the optional 8-bit copy are not stored: ``b200rag_load`` rebuilds them on the device."""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

MAGIC = b"B200RAG1"
HEADER = struct.Struct("<8s4i2q")
assert HEADER.size == 40


@dataclass
class ShardFile:
    version: int
    dim: int
    vocab: int
    n_rows: int
    nnz: int
    dense: np.ndarray            # uint16 [n_rows, dim]
    indptr: np.ndarray           # int64 [n_rows + 1]
    terms: np.ndarray            # uint32 [nnz]
    weights: np.ndarray          # float32 [nnz]
    row_ids: np.ndarray | None   # int64 [n_rows]; None in version 1 files


def read_header(path: str):
    with open(path, "rb") as f:
        raw = f.read(HEADER.size)
    if len(raw) != HEADER.size:
        raise ValueError(f"{path}: not a b200rag shard file (short header)")
    magic, version, dim, vocab, _reserved, n_rows, nnz = HEADER.unpack(raw)
    if magic != MAGIC or version not in (1, 2):
        raise ValueError(f"{path}: not a b200rag shard file")
    if dim <= 0 or vocab <= 0 or n_rows < 0 or nnz < 0:
        raise ValueError(f"{path}: corrupt header")
    return version, dim, vocab, n_rows, nnz


def expected_size(version: int, dim: int, n_rows: int, nnz: int) -> int:
    return HEADER.size + n_rows * dim * 2 + (n_rows + 1) * 8 + nnz * 8 + (n_rows * 8 if version >= 2 else 0)


def open_mmap(path: str, mode: str = "r") -> ShardFile:
    """Map a shard file's arrays in place (numpy.memmap views; nothing is read until it is touched)."""
    version, dim, vocab, n_rows, nnz = read_header(path)
    size = os.path.getsize(path)
    if size != expected_size(version, dim, n_rows, nnz):
        raise ValueError(f"{path}: {size} bytes, the header promises {expected_size(version, dim, n_rows, nnz)}")

    def view(dtype, shape, offset):
        count = int(np.prod(shape))
        if count == 0:
            return np.zeros(shape, dtype=dtype)
        return np.memmap(path, dtype=dtype, mode=mode, offset=offset, shape=shape)

    off = HEADER.size
    dense = view(np.uint16, (n_rows, dim), off)
    off += n_rows * dim * 2
    indptr = view(np.int64, (n_rows + 1,), off)
    off += (n_rows + 1) * 8
    terms = view(np.uint32, (nnz,), off)
    off += nnz * 4
    weights = view(np.float32, (nnz,), off)
    off += nnz * 4
    row_ids = view(np.int64, (n_rows,), off) if version >= 2 else None
    if int(indptr[0]) != 0 or int(indptr[n_rows]) != nnz:
        raise ValueError(f"{path}: forward index does not span the postings")
    return ShardFile(version, dim, vocab, n_rows, nnz, dense, indptr, terms, weights, row_ids)


def write(path: str, dense_bits, indptr, terms, weights, row_ids=None, *, vocab: int = 250_002, row_base: int = 0):
    """Write a version-2 shard file from host arrays (what ``b200rag_save`` writes for the same rows).

    dense_bits: uint16 [n, dim] bf16 bits of UNIT rows (``b200rag.normalize_bf16``); indptr / terms / weights: the rows'
    sparse vectors in CSR form, terms ascending and unique per row; row_ids: strictly increasing global ids
    (default ``row_base + arange(n)``).  The file is written to a temporary name and renamed into place."""
    dense_bits = np.ascontiguousarray(dense_bits, dtype=np.uint16)
    if dense_bits.ndim != 2:
        raise ValueError("dense_bits must be [n_rows, dim]")
    n, dim = dense_bits.shape
    indptr = np.ascontiguousarray(indptr if indptr is not None else np.zeros(n + 1), dtype=np.int64)
    terms = np.ascontiguousarray(terms if terms is not None else np.zeros(0), dtype=np.uint32)
    weights = np.ascontiguousarray(weights if weights is not None else np.zeros(0), dtype=np.float32)
    if len(indptr) != n + 1 or indptr[0] != 0 or (np.diff(indptr) < 0).any() or indptr[n] != len(terms) or len(terms) != len(weights):
        raise ValueError("indptr / terms / weights are not a CSR over the rows")
    if len(terms) and int(terms.max()) >= vocab:
        raise ValueError("term id outside the vocabulary")
    if len(terms) > 1:                                   # b200rag_load trusts this (include/b200rag.h): check it here
        asc = np.diff(terms.astype(np.int64)) > 0
        starts = indptr[1:-1]
        asc[starts[(starts > 0) & (starts < len(terms))] - 1] = True      # a new row may start with any term
        if not asc.all():
            raise ValueError("terms must be ascending and unique within every row")
    ids = np.arange(row_base, row_base + n, dtype=np.int64) if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.int64)
    if len(ids) != n or (n > 1 and (np.diff(ids) <= 0).any()):
        raise ValueError("row ids must be strictly increasing, one per row")
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(HEADER.pack(MAGIC, 2, dim, vocab, 0, n, len(terms)))
        for a in (dense_bits, indptr, terms, weights, ids):
            a.tofile(f)
    os.replace(tmp, path)
