"""ctypes binding of libb200rag.so (include/b200rag.h).

This is the FFI a maintainer of the reference would bind in place of ``qdrant_client.QdrantClient``
inside ``QdrantRetriever`` (/root/reference/src/audio_rag/retrieval/qdrant.py:35-54, 197-220, 281-332).
There is no CPU fallback: importing works without a GPU (so the host logic and the symbol table can be
tested), but every compute call raises ``B200RagError`` when no sm_100 device is present, and a missing
library is an ImportError, never a silent downgrade.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200rag.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOGPU, ERR_OOM, ERR_STATE, ERR_INEXACT = range(7)
DENSE, SPARSE, HYBRID = 0, 1, 2
MODES = {"dense": DENSE, "sparse": SPARSE, "hybrid": HYBRID}
MAX_TOPK = 256


class B200RagError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200rag error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("dim", C.c_int32), ("vocab", C.c_int32), ("docs_per_block", C.c_int32),
                ("row_base", C.c_int64), ("reserve_rows", C.c_int64), ("reserve_postings", C.c_int64)]


class Cand(C.Structure):
    _fields_ = [("id", C.c_int64), ("score", C.c_float), ("valid", C.c_uint32)]


CAND_DTYPE = np.dtype([("id", np.int64), ("score", np.float32), ("valid", np.uint32)])


class Query(C.Structure):
    _fields_ = [("mode", C.c_int32), ("batch", C.c_int32), ("top_k", C.c_int32), ("rrf_k", C.c_int32),
                ("has_threshold", C.c_int32), ("score_threshold", C.c_float),
                ("q_dense_bits", C.c_void_p), ("q_sp_indptr", C.c_void_p), ("q_sp_terms", C.c_void_p),
                ("q_sp_weights", C.c_void_p), ("mask_ids", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int32), ("dense_path", C.c_int32), ("dense_bytes", C.c_int64),
                ("sparse_postings", C.c_int64), ("dense_passes", C.c_int32), ("retries", C.c_int32),
                ("dense_scan_ms", C.c_float), ("sparse_scan_ms", C.c_float),
                ("pre_scan_ms", C.c_float), ("tail_ms", C.c_float), ("exhaustive", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/b200rag.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("b200rag_version", C.c_char_p, []),
    ("b200rag_last_error", C.c_char_p, []),
    ("b200rag_device_count", C.c_int, []),
    ("b200rag_normalize_bf16", C.c_int, [_P, C.c_int64, C.c_int32, _P]),
    ("b200rag_shard_create", C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    ("b200rag_shard_destroy", None, [_P]),
    ("b200rag_set_stream", C.c_int, [_P, _P]),
    ("b200rag_set_slack", C.c_int, [_P, C.c_int32]),
    ("b200rag_set_exhaustive", C.c_int, [_P, C.c_int32]),
    ("b200rag_set_compression", C.c_int, [_P, C.c_int32]),
    ("b200rag_set_exact_fallback", C.c_int, [_P, C.c_int32]),
    ("b200rag_set_pipeline", C.c_int, [_P, C.c_int32, _P]),
    ("b200rag_result_stream", C.c_void_p, [_P]),
    ("b200rag_pipeline_pause", C.c_int, [_P, C.c_int32]),
    ("b200rag_set_dense_path", C.c_int, [_P, C.c_int32]),
    ("b200rag_debug_dense_scores", C.c_int, [_P, _P]),
    ("b200rag_sync", C.c_int, [_P]),
    ("b200rag_add", C.c_int, [_P, C.c_int64, _P, _P, _P, _P]),
    ("b200rag_add_device", C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_int64]),
    ("b200rag_add_ids", C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P]),
    ("b200rag_add_device_ids", C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_int64, _P]),
    ("b200rag_add_f32", C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P]),
    ("b200rag_compact", C.c_int, [_P, _P, C.c_int64]),
    ("b200rag_build", C.c_int, [_P]),
    ("b200rag_count", C.c_int64, [_P]),
    ("b200rag_postings", C.c_int64, [_P]),
    ("b200rag_clear", C.c_int, [_P]),
    ("b200rag_read_dense", C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    ("b200rag_read_sparse", C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, C.c_int64]),
    ("b200rag_read_row_ids", C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    ("b200rag_mask_set", C.c_int, [_P, C.c_int32, _P, C.c_int64]),
    ("b200rag_mask_set_device", C.c_int, [_P, C.c_int32, _P, C.c_int64]),
    ("b200rag_mask_drop", C.c_int, [_P, C.c_int32]),
    ("b200rag_search", C.c_int, [_P, C.POINTER(Query), _P, _P, _P]),
    ("b200rag_stage", C.c_int, [_P, C.POINTER(Query)]),
    ("b200rag_stage_slot", C.c_int, [_P, C.POINTER(Query), C.c_int32]),
    ("b200rag_use_slot", C.c_int, [_P, C.c_int32]),
    ("b200rag_normalize_bf16_device", C.c_int, [_P, _P, C.c_int64, _P]),
    ("b200rag_stage_device", C.c_int, [_P, C.POINTER(Query), C.c_int32]),
    ("b200rag_legs_len", C.c_int, [C.POINTER(Query), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("b200rag_legs", C.c_int, [_P, _P, _P]),
    ("b200rag_fuse", C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    ("b200rag_save", C.c_int, [_P, C.c_char_p]),
    ("b200rag_load", C.c_int, [_P, C.c_char_p]),
    ("b200rag_p2p_export", C.c_int, [_P, C.c_int32, C.c_int64, _P]),
    ("b200rag_p2p_attach", C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    ("b200rag_p2p_exchange", C.c_int, [_P, _P, C.c_int64]),
    ("b200rag_p2p_fuse", C.c_int, [_P, _P, _P, _P]),
    ("b200rag_p2p_close", C.c_int, [_P]),
    ("b200rag_p2p_set_stream", C.c_int, [_P, _P]),
    ("b200rag_group_create", C.c_int, [C.POINTER(_P), C.c_int32, C.POINTER(_P)]),
    ("b200rag_group_destroy", None, [_P]),
    ("b200rag_group_size", C.c_int32, [_P]),
    ("b200rag_group_search", C.c_int, [_P, C.POINTER(Query), _P, _P, _P]),
    ("b200rag_group_get_stats", C.c_int, [_P, C.POINTER(Stats)]),
    ("b200rag_get_stats", C.c_int, [_P, C.POINTER(Stats)]),
    ("b200rag_get_stats_step", C.c_int, [_P, C.c_int32, C.POINTER(Stats)]),
    ("b200rag_set_profiling", C.c_int, [_P, C.c_int32]),
    ("b200rag_synth_dense", C.c_int, [_P, C.c_uint64, C.c_int64, C.c_int64, _P]),
    ("b200rag_synth_sparse", C.c_int, [_P, C.c_uint64, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P, C.c_int64,
                                       _P, _P, _P, _P]),
    ("b200rag_exclusive_scan_i64", C.c_int, [_P, _P, C.c_int64, _P]),
    ("b200rag_synth_collection_mask", C.c_int, [_P, C.c_uint64, C.c_int64, C.c_int64, _P, C.c_int32, C.c_int32, _P]),
]

_lib = None


def load():
    """Load the shared library (ImportError if it was not built: there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python audio-rag_b200/build.py` "
                              "(or __graft_entry__.build()); b200rag has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != OK:
        raise B200RagError(rc, load().b200rag_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return int(load().b200rag_device_count())


def _np_ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _ptr(x):
    """numpy array -> host pointer, int -> raw (device) pointer, object with data_ptr() -> device pointer."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))


def normalize_bf16(x: np.ndarray) -> np.ndarray:
    """Host routine of the library (no GPU): unit-normalise fp32 rows and round to bf16 bits."""
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    check(load().b200rag_normalize_bf16(_np_ptr(x), x.shape[0], x.shape[1], _np_ptr(out)))
    return out


class Shard:
    """One row-range shard of the corpus on one GPU (thin, typed wrapper over the C ABI)."""

    def __init__(self, dim=1024, vocab=250_002, device=0, row_base=0, docs_per_block=0, reserve_rows=0,
                 reserve_postings=0):
        self._lib = load()
        self.dim, self.vocab, self.device, self.row_base = dim, vocab, device, row_base
        cfg = Config(device, dim, vocab, docs_per_block, row_base, reserve_rows, reserve_postings)
        h = C.c_void_p()
        check(self._lib.b200rag_shard_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200rag_shard_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ingest
    @staticmethod
    def _csr(n, sp_indptr, sp_terms, sp_weights, ids):
        if sp_indptr is not None:
            sp_indptr = np.ascontiguousarray(sp_indptr, dtype=np.int64)
            sp_terms = np.ascontiguousarray(sp_terms, dtype=np.uint32)
            sp_weights = np.ascontiguousarray(sp_weights, dtype=np.float32)
            if len(sp_indptr) != n + 1:
                raise B200RagError(ERR_INVALID, "sparse indptr length must be n+1")
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            if len(ids) != n:
                raise B200RagError(ERR_INVALID, "one row id per row expected")
        return sp_indptr, sp_terms, sp_weights, ids

    def add(self, dense_bits: np.ndarray, sp_indptr=None, sp_terms=None, sp_weights=None, ids=None):
        """Append unit bf16 rows (+ their sparse parts); `ids` = global row ids (None: row_base + local row)."""
        dense_bits = np.ascontiguousarray(dense_bits, dtype=np.uint16).reshape(-1, self.dim)
        n = dense_bits.shape[0]
        sp_indptr, sp_terms, sp_weights, ids = self._csr(n, sp_indptr, sp_terms, sp_weights, ids)
        check(self._lib.b200rag_add_ids(self._h, n, _np_ptr(dense_bits), _np_ptr(sp_indptr), _np_ptr(sp_terms),
                                        _np_ptr(sp_weights), _np_ptr(ids)))

    def add_f32(self, dense_f32: np.ndarray, sp_indptr=None, sp_terms=None, sp_weights=None, ids=None):
        """Append RAW fp32 rows: normalised + rounded to bf16 on the GPU (bit-equal to normalize_bf16 + add)."""
        dense_f32 = np.ascontiguousarray(dense_f32, dtype=np.float32).reshape(-1, self.dim)
        n = dense_f32.shape[0]
        sp_indptr, sp_terms, sp_weights, ids = self._csr(n, sp_indptr, sp_terms, sp_weights, ids)
        check(self._lib.b200rag_add_f32(self._h, n, _np_ptr(dense_f32), _np_ptr(sp_indptr), _np_ptr(sp_terms),
                                        _np_ptr(sp_weights), _np_ptr(ids)))

    def add_device(self, n, dense_bits_dev, sp_indptr_dev=None, sp_terms_dev=None, sp_weights_dev=None, nnz=0, ids=None):
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
        check(self._lib.b200rag_add_device_ids(self._h, n, _ptr(dense_bits_dev), _ptr(sp_indptr_dev),
                                               _ptr(sp_terms_dev), _ptr(sp_weights_dev), nnz, _np_ptr(ids)))

    def compact(self, keep_words: np.ndarray, n_rows: int):
        """Physically drop the local rows whose keep bit is clear (ids and order of the others are preserved)."""
        keep_words = np.ascontiguousarray(keep_words, dtype=np.uint32)
        check(self._lib.b200rag_compact(self._h, _np_ptr(keep_words), n_rows))

    def build(self):
        check(self._lib.b200rag_build(self._h))

    def clear(self):
        check(self._lib.b200rag_clear(self._h))

    @property
    def count(self) -> int:
        return int(self._lib.b200rag_count(self._h))

    @property
    def postings(self) -> int:
        return int(self._lib.b200rag_postings(self._h))

    def read_dense(self, row: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), dtype=np.uint16)
        check(self._lib.b200rag_read_dense(self._h, row, n, _np_ptr(out)))
        return out

    def read_sparse(self, row: int, n: int):
        """(indptr int64[n+1] from 0, terms uint32, weights float32) of the stored rows [row, row + n)."""
        indptr = np.empty(n + 1, dtype=np.int64)
        check(self._lib.b200rag_read_sparse(self._h, row, n, _np_ptr(indptr), None, None, 0))
        nnz = int(indptr[n])
        terms = np.empty(nnz, dtype=np.uint32)
        w = np.empty(nnz, dtype=np.float32)
        if nnz:
            check(self._lib.b200rag_read_sparse(self._h, row, n, _np_ptr(indptr), _np_ptr(terms), _np_ptr(w), nnz))
        return indptr, terms, w

    def read_row_ids(self, row: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int64)
        check(self._lib.b200rag_read_row_ids(self._h, row, n, _np_ptr(out)))
        return out

    # ---- masks
    def mask_set(self, mask_id: int, words, n_rows: int):
        if isinstance(words, np.ndarray):
            words = np.ascontiguousarray(words, dtype=np.uint32)
            check(self._lib.b200rag_mask_set(self._h, mask_id, _np_ptr(words), n_rows))
        else:
            check(self._lib.b200rag_mask_set_device(self._h, mask_id, _ptr(words), n_rows))

    def mask_drop(self, mask_id: int):
        check(self._lib.b200rag_mask_drop(self._h, mask_id))

    # ---- search
    def set_stream(self, stream_ptr):
        """stream_ptr: a cudaStream_t as int (0/None = the legacy default stream, torch's default current stream)."""
        check(self._lib.b200rag_set_stream(self._h, C.c_void_p(stream_ptr) if stream_ptr else None))

    def set_slack(self, slack: int):
        check(self._lib.b200rag_set_slack(self._h, slack))

    def set_exhaustive(self, on: bool):
        """Legs score EVERY eligible row canonically and sort: always exact (fallback + cross-check path)."""
        check(self._lib.b200rag_set_exhaustive(self._h, 1 if on else 0))

    def set_compression(self, on: bool):
        """Opt-in 8-bit copy of the rows for the candidate scan of 1-2 query searches (results stay exact)."""
        check(self._lib.b200rag_set_compression(self._h, 1 if on else 0))

    def set_exact_fallback(self, on: bool):
        """Off: a search whose slack guard never clears raises B200RagError(ERR_INEXACT) instead of falling back."""
        check(self._lib.b200rag_set_exact_fallback(self._h, 1 if on else 0))

    def set_pipeline(self, on: bool, stream_ptr: int = 0):
        """Throughput mode for back-to-back staged searches: only the dense scan stays on the shard's stream, the tails,
        the exchange and the fuse run on `result_stream()` -- `stream_ptr` (a cudaStream_t of the caller) or a stream of
        the library's (see include/b200rag.h)."""
        check(self._lib.b200rag_set_pipeline(self._h, 1 if on else 0, C.c_void_p(stream_ptr) if stream_ptr else None))

    def pipeline_pause(self, on: bool):
        """Classic-form searches on a pipelined shard (no CUDA call; see include/b200rag.h)."""
        check(self._lib.b200rag_pipeline_pause(self._h, 1 if on else 0))

    def result_stream(self) -> int:
        """cudaStream_t (as int) on which a search's fused results become available."""
        return int(self._lib.b200rag_result_stream(self._h) or 0)

    def set_dense_path(self, path: int):
        """0 = auto, 1 = SIMT bulk-copy scan, 2 = tcgen05 GEMM."""
        check(self._lib.b200rag_set_dense_path(self._h, path))

    def debug_dense_scores(self, out_scores_dev):
        check(self._lib.b200rag_debug_dense_scores(self._h, _ptr(out_scores_dev)))

    def sync(self):
        check(self._lib.b200rag_sync(self._h))

    def make_query(self, mode, top_k, q_bits=None, sp_indptr=None, sp_terms=None, sp_weights=None, mask_ids=None,
                   score_threshold=None, rrf_k=0):
        """Build the C query struct; returns (struct, keepalive list)."""
        mode = MODES[mode] if isinstance(mode, str) else int(mode)
        keep = []
        batch = None
        if q_bits is not None:
            q_bits = np.ascontiguousarray(q_bits, dtype=np.uint16).reshape(-1, self.dim)
            batch = q_bits.shape[0]
            keep.append(q_bits)
        if sp_indptr is not None:
            sp_indptr = np.ascontiguousarray(sp_indptr, dtype=np.int64)
            sp_terms = np.ascontiguousarray(sp_terms, dtype=np.uint32)
            sp_weights = np.ascontiguousarray(sp_weights, dtype=np.float32)
            keep += [sp_indptr, sp_terms, sp_weights]
            batch = len(sp_indptr) - 1 if batch is None else batch
        if mask_ids is not None:
            mask_ids = np.ascontiguousarray(mask_ids, dtype=np.int32)
            keep.append(mask_ids)
        q = Query(mode, batch or 0, top_k, rrf_k, 0 if score_threshold is None else 1,
                  0.0 if score_threshold is None else float(score_threshold),
                  q_bits.ctypes.data if q_bits is not None else None,
                  sp_indptr.ctypes.data if sp_indptr is not None else None,
                  sp_terms.ctypes.data if sp_terms is not None else None,
                  sp_weights.ctypes.data if sp_weights is not None else None,
                  mask_ids.ctypes.data if mask_ids is not None else None)
        return q, keep

    def search(self, mode, top_k, q_bits=None, sp_indptr=None, sp_terms=None, sp_weights=None, mask_ids=None,
               score_threshold=None, rrf_k=0):
        """Whole path, host buffers in and out.  Returns (ids [B,k] int64, scores [B,k] float64, counts [B])."""
        q, keep = self.make_query(mode, top_k, q_bits, sp_indptr, sp_terms, sp_weights, mask_ids, score_threshold,
                                  rrf_k)
        ids = np.empty((q.batch, top_k), dtype=np.int64)
        scores = np.empty((q.batch, top_k), dtype=np.float64)
        counts = np.empty(q.batch, dtype=np.int32)
        check(self._lib.b200rag_search(self._h, C.byref(q), _np_ptr(ids), _np_ptr(scores), _np_ptr(counts)))
        del keep
        return ids, scores, counts

    def stage(self, q: Query, keep=None, slot: int = 0):
        """Copy a query batch to the device (its own block per `slot`) and make it the active one."""
        self._keep = keep
        check(self._lib.b200rag_stage_slot(self._h, C.byref(q), slot))

    def normalize_bf16_device(self, x_dev, n: int, out_bits_dev):
        """Device twin of normalize_bf16 (bit-equal): fp32 rows [n, dim] -> unit bf16 bits [n, dim], both on the GPU."""
        check(self._lib.b200rag_normalize_bf16_device(self._h, _ptr(x_dev), n, _ptr(out_bits_dev)))

    def stage_device(self, mode, top_k, batch, q_bits_dev=None, sp_indptr=None, sp_terms_dev=None, sp_weights_dev=None,
                     mask_ids=None, score_threshold=None, rrf_k=0, slot: int = 0):
        """Stage a batch whose vectors / sparse terms / weights already live on the GPU (device pointers or tensors);
        `sp_indptr` (batch + 1) and `mask_ids` are small host arrays.  Returns the query struct (for legs_len)."""
        mode = MODES[mode] if isinstance(mode, str) else int(mode)
        keep = []
        if sp_indptr is not None:
            sp_indptr = np.ascontiguousarray(sp_indptr, dtype=np.int64)
            keep.append(sp_indptr)
        if mask_ids is not None:
            mask_ids = np.ascontiguousarray(mask_ids, dtype=np.int32)
            keep.append(mask_ids)

        def dp(x):
            p = _ptr(x)
            return None if p is None else p.value

        q = Query(mode, batch, top_k, rrf_k, 0 if score_threshold is None else 1,
                  0.0 if score_threshold is None else float(score_threshold), dp(q_bits_dev),
                  sp_indptr.ctypes.data if sp_indptr is not None else None, dp(sp_terms_dev), dp(sp_weights_dev),
                  mask_ids.ctypes.data if mask_ids is not None else None)
        self._keep = keep
        check(self._lib.b200rag_stage_device(self._h, C.byref(q), slot))
        return q

    def use_slot(self, slot: int):
        """Re-activate an already staged batch (no copy, no synchronisation)."""
        check(self._lib.b200rag_use_slot(self._h, slot))

    @staticmethod
    def legs_len(q: Query):
        nlegs, L = C.c_int32(), C.c_int32()
        check(load().b200rag_legs_len(C.byref(q), C.byref(nlegs), C.byref(L)))
        return nlegs.value, L.value

    def legs(self, cands_dev, ambiguous_dev=None):
        check(self._lib.b200rag_legs(self._h, _ptr(cands_dev), _ptr(ambiguous_dev)))

    def fuse(self, gathered_dev, n_shards, out_ids_dev, out_scores_dev, out_counts_dev, has_trailer=False):
        check(self._lib.b200rag_fuse(self._h, _ptr(gathered_dev), n_shards, 1 if has_trailer else 0, _ptr(out_ids_dev),
                                     _ptr(out_scores_dev), _ptr(out_counts_dev)))

    def save(self, path: str):
        check(self._lib.b200rag_save(self._h, os.fsencode(path)))

    def load(self, path: str):
        check(self._lib.b200rag_load(self._h, os.fsencode(path)))

    # ---- peer-memory candidate exchange (replaces the all-gather between legs and fuse on one box)
    IPC_HANDLE_BYTES = 64

    def p2p_export(self, world: int, slot_bytes: int) -> bytes:
        h = (C.c_uint8 * self.IPC_HANDLE_BYTES)()
        check(self._lib.b200rag_p2p_export(self._h, world, slot_bytes, C.cast(h, _P)))
        return bytes(h)

    def p2p_attach(self, rank: int, world: int, handles: bytes):
        assert len(handles) == world * self.IPC_HANDLE_BYTES
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        check(self._lib.b200rag_p2p_attach(self._h, rank, world, C.cast(buf, _P)))

    def p2p_exchange(self, mine_dev, nbytes: int):
        check(self._lib.b200rag_p2p_exchange(self._h, _ptr(mine_dev), nbytes))

    def p2p_fuse(self, out_ids_dev, out_scores_dev, out_counts_dev):
        check(self._lib.b200rag_p2p_fuse(self._h, _ptr(out_ids_dev), _ptr(out_scores_dev), _ptr(out_counts_dev)))

    def p2p_set_stream(self, stream_ptr):
        check(self._lib.b200rag_p2p_set_stream(self._h, C.c_void_p(stream_ptr) if stream_ptr else None))

    def p2p_close(self):
        check(self._lib.b200rag_p2p_close(self._h))

    def set_profiling(self, on: bool):
        check(self._lib.b200rag_set_profiling(self._h, 1 if on else 0))

    def stats(self) -> dict:
        st = Stats()
        check(self._lib.b200rag_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in Stats._fields_}

    def stats_step(self, steps_back: int) -> dict:
        """Event timings (the *_ms fields) of the legs call made `steps_back` calls ago; profiling must be on."""
        st = Stats()
        check(self._lib.b200rag_get_stats_step(self._h, steps_back, C.byref(st)))
        return {k: getattr(st, k) for k in ("dense_scan_ms", "sparse_scan_ms", "pre_scan_ms", "tail_ms")}

    # ---- synthetic generation on the device
    def synth_dense(self, seed, global_row_start, n, out_dev):
        check(self._lib.b200rag_synth_dense(self._h, seed, global_row_start, n, _ptr(out_dev)))

    def synth_sparse(self, seed, global_row_start, n, doc_tokens, thr_dev, idf_dev, tff_dev, term_mul, counts_dev,
                     indptr_dev, terms_dev, weights_dev):
        check(self._lib.b200rag_synth_sparse(self._h, seed, global_row_start, n, doc_tokens, _ptr(thr_dev),
                                             _ptr(idf_dev), _ptr(tff_dev), term_mul, _ptr(counts_dev),
                                             _ptr(indptr_dev), _ptr(terms_dev), _ptr(weights_dev)))

    def exclusive_scan_i64(self, in_dev, n, out_dev):
        check(self._lib.b200rag_exclusive_scan_i64(self._h, _ptr(in_dev), n, _ptr(out_dev)))

    def synth_collection_mask(self, seed, global_row_start, n, thr_dev, n_collections, collection, out_words_dev):
        check(self._lib.b200rag_synth_collection_mask(self._h, seed, global_row_start, n, _ptr(thr_dev),
                                                      n_collections, collection, _ptr(out_words_dev)))


class ShardGroup:
    """Several shards of one corpus driven by this process (b200rag_group_*): `search` is `Shard.search` over all of
    them, bit-identical to one shard holding every row.  The shards stay owned by the caller."""

    def __init__(self, shards):
        self._lib = load()
        self.shards = list(shards)
        self.dim = self.shards[0].dim
        arr = (_P * len(self.shards))(*[s._h for s in self.shards])
        h = C.c_void_p()
        check(self._lib.b200rag_group_create(arr, len(self.shards), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200rag_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, mode, top_k, q_bits=None, sp_indptr=None, sp_terms=None, sp_weights=None, mask_ids=None,
               score_threshold=None, rrf_k=0):
        q, keep = self.shards[0].make_query(mode, top_k, q_bits, sp_indptr, sp_terms, sp_weights, mask_ids,
                                            score_threshold, rrf_k)
        ids = np.empty((q.batch, top_k), dtype=np.int64)
        scores = np.empty((q.batch, top_k), dtype=np.float64)
        counts = np.empty(q.batch, dtype=np.int32)
        check(self._lib.b200rag_group_search(self._h, C.byref(q), _np_ptr(ids), _np_ptr(scores), _np_ptr(counts)))
        del keep
        return ids, scores, counts

    def stats(self) -> dict:
        st = Stats()
        check(self._lib.b200rag_group_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in Stats._fields_}
