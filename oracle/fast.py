"""ctypes binding of oracle/oracle_c.c (TEST INFRASTRUCTURE ONLY; see oracle/oracle.py header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def dense_scores(bits, q_bits):
    bits = np.ascontiguousarray(bits, dtype=np.uint16)
    q = np.ascontiguousarray(q_bits, dtype=np.uint16)
    n, dim = bits.shape
    out = np.empty(n, dtype=np.float32)
    lib().oracle_dense_scores(_p(bits, C.c_uint16), C.c_int64(n), C.c_int32(dim), _p(q, C.c_uint16),
                              _p(out, C.c_float))
    return out


def sparse_scores(indptr, terms, weights, q_idx, q_val):
    from .oracle import check_sparse_vector
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    terms = np.ascontiguousarray(terms, dtype=np.uint32)
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    qi, qv = check_sparse_vector(q_idx, q_val)
    qi = np.ascontiguousarray(qi, dtype=np.int64)
    qv = np.ascontiguousarray(qv, dtype=np.float32)
    n = len(indptr) - 1
    out = np.empty(n, dtype=np.float32)
    touched = np.empty(n, dtype=np.uint8)
    lib().oracle_sparse_scores(_p(indptr, C.c_int64), _p(terms, C.c_uint32), _p(weights, C.c_float),
                               C.c_int64(n), _p(qi, C.c_int64), _p(qv, C.c_float), C.c_int32(len(qi)),
                               _p(out, C.c_float), _p(touched, C.c_uint8))
    return out, touched.astype(bool)


def synth_dense_bf16(seed, row_start, n, dim=1024):
    out = np.empty((n, dim), dtype=np.uint16)
    lib().oracle_synth_dense_bf16(C.c_uint64(seed), C.c_int64(row_start), C.c_int64(n), C.c_int32(dim),
                                  _p(out, C.c_uint16))
    return out


def synth_sparse_csr(seed, row_start, n, thresholds, idf, tff, vocab, doc_tokens, term_mul):
    thr = np.ascontiguousarray(thresholds, dtype=np.uint64)
    idf = np.ascontiguousarray(idf, dtype=np.float32)
    tff = np.ascontiguousarray(tff, dtype=np.float32)
    counts = np.zeros(n, dtype=np.int64)
    f = lib().oracle_synth_sparse
    args = [C.c_uint64(seed), C.c_int64(row_start), C.c_int64(n), C.c_int32(vocab), C.c_int32(doc_tokens),
            _p(thr, C.c_uint64), _p(idf, C.c_float), _p(tff, C.c_float), C.c_int64(term_mul)]
    f(*args, _p(counts, C.c_int64), None, None, None)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    terms = np.empty(int(indptr[-1]), dtype=np.uint32)
    w = np.empty(int(indptr[-1]), dtype=np.float32)
    f(*args, None, _p(indptr, C.c_int64), _p(terms, C.c_uint32), _p(w, C.c_float))
    return indptr, terms, w
