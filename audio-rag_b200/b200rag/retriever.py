"""B200Retriever: the reference's retrieval plugin interface on top of the CUDA shard engine.

Mirror of ``QdrantRetriever`` (/root/reference/src/audio_rag/retrieval/qdrant.py:14-381): same constructor,
same ``add / search / count / collection_exists / delete_collection / is_hybrid_collection`` surface, same
defaults (``collection_name or config.collection_name`` :56-57, ``top_k or config.top_k`` :249,
``search_type or config.search_type`` :250), same branch/fallback rules (:272, :299, :313), the same
``limit = 2 * top_k`` per hybrid leg (:287, :292) and the same error convention (every failure surfaces as
``RetrievalError``).  What the reference delegates to qdrant-client (scoring, candidate selection, RRF) runs
in libb200rag.so; what stays on the host is what is host work in the reference too: the payload store,
the collection registry and result materialisation (:334-346).

All collections share one global row space (row id = insertion order; ids are never reused or
renumbered); a collection, its tombstones and an optional ``filter_metadata`` are expressed as one
eligibility bitmask per search (north star: "the collection filter applied as a bitmask"), applied in BOTH
legs.

Several GPUs behind the ONE retriever object the reference builds (pipeline/orchestrator.py:48-74):
``devices=[0, 1, ...]`` (or the environment variable ``B200RAG_DEVICES=all | 0,1,2``) makes the retriever own
one shard per entry, all in this process.  ``add()`` routes each batch to a shard together with its global row
ids, masks are cut per shard, and ``search`` / ``search_batch`` run the legs on every shard and fuse them
(``b200rag_group_search``); the results are bit-identical to a single shard holding every row.

Additive API (no reference counterpart): ``search_batch``, ``search_batch_arrays``, ``save`` / ``load``,
``attach_prebuilt``.
"""
from __future__ import annotations

import struct
import json
import os
from collections import OrderedDict

import numpy as np

from . import _ffi
from .compat import (AudioChunk, BaseRetriever, EmbeddingResult, RetrievalConfig, RetrievalError, RetrievalResult,
                     get_logger, timed)
from .synth import pack_mask

logger = get_logger(__name__)

try:  # register next to "qdrant" when the reference package is importable (core/registry.py:27-34)
    from audio_rag.retrieval.base import RetrievalRegistry  # type: ignore

    def _register(cls):
        if "b200" not in RetrievalRegistry:
            return RetrievalRegistry.register("b200")(cls)
        return cls
except Exception:  # reference not importable here
    def _register(cls):
        return cls

_MASK_CACHE = 64
_PAYLOAD_CHUNK = 65536          # payloads per file in save()


def _sorted_sparse(sv, vocab):
    """indices/values lists -> (uint32 ascending, float32); duplicates and out-of-range indices are rejected (R3)."""
    ind, vals = sv.indices, sv.values
    if isinstance(ind, list) and isinstance(vals, list) and len(ind) <= 64:
        # a query's handful of terms (5-30 typical): plain Python beats six numpy calls on a dozen items
        n = len(ind)
        if n != len(vals):
            raise RetrievalError(f"sparse vector has {n} indices but {len(vals)} values")
        if n == 0:
            return np.zeros(0, np.uint32), np.zeros(0, np.float32)
        try:
            pairs = sorted(zip(ind, vals))
            lo, hi = pairs[0][0], pairs[-1][0]
            if type(lo) is not int or type(hi) is not int:
                raise TypeError
        except TypeError:
            pairs = None                                  # odd element types: let numpy decide, as before
        if pairs is not None:
            if lo < 0 or hi >= vocab:
                raise RetrievalError(f"sparse index out of range [0, {vocab})")
            idx = [p[0] for p in pairs]
            for a, b in zip(idx, idx[1:]):
                if a == b:
                    raise RetrievalError("duplicate index in sparse vector")
            return np.array(idx, dtype=np.uint32), np.array([p[1] for p in pairs], dtype=np.float32)
    idx = np.asarray(sv.indices, dtype=np.int64).reshape(-1)
    val = np.asarray(sv.values, dtype=np.float32).reshape(-1)
    if idx.shape != val.shape:
        raise RetrievalError(f"sparse vector has {len(idx)} indices but {len(val)} values")
    if len(idx):
        if idx.min() < 0 or idx.max() >= vocab:
            raise RetrievalError(f"sparse index out of range [0, {vocab})")
        o = np.argsort(idx, kind="stable")
        idx, val = idx[o], val[o]
        if (idx[1:] == idx[:-1]).any():
            raise RetrievalError("duplicate index in sparse vector")
    return idx.astype(np.uint32), val


def _match(payload_meta, key, value):
    """FieldCondition(key="metadata.<k>", match=MatchValue(value=v)) (qdrant.py:264-268): equality, or membership
    when the stored value is a list."""
    if not isinstance(payload_meta, dict) or key not in payload_meta:
        return False
    got = payload_meta[key]
    if isinstance(got, (list, tuple)):
        return value in got
    return got == value


class _Grow:
    """Append-only numpy vector that grows by doubling (row bookkeeping scales with the corpus, not with objects)."""

    def __init__(self, dtype, cap=1024):
        self._buf = np.zeros(cap, dtype=dtype)
        self.n = 0

    @property
    def view(self):
        return self._buf[:self.n]

    def reserve(self, extra):
        need = self.n + extra
        if need > len(self._buf):
            cap = len(self._buf)
            while cap < need:
                cap *= 2
            grown = np.zeros(cap, dtype=self._buf.dtype)
            grown[:self.n] = self._buf[:self.n]
            self._buf = grown

    def append(self, values):
        values = np.asarray(values, dtype=self._buf.dtype).reshape(-1)
        self.reserve(len(values))
        self._buf[self.n:self.n + len(values)] = values
        self.n += len(values)

    def fill(self, count, value):
        self.reserve(count)
        self._buf[self.n:self.n + count] = value
        self.n += count

    def assign(self, values):
        values = np.asarray(values, dtype=self._buf.dtype).reshape(-1)
        self.n = 0
        self.append(values)


class _PayloadStore:
    """Payload per global row id.  A list, optionally preceded by a LAZY range whose payloads come from a callable
    (``attach_prebuilt``: a 10M-row synthetic corpus needs no 10M Python dicts)."""

    def __init__(self):
        self._lazy_n, self._lazy_fn, self._items = 0, None, []

    def __len__(self):
        return self._lazy_n + len(self._items)

    def __getitem__(self, i):
        j = i - self._lazy_n
        if j >= 0:                                  # (the per-hit lookup of every search: no len() calls)
            if j >= len(self._items):
                raise IndexError(i)
            return self._items[j]
        if i < 0:
            raise IndexError(i)
        return self._lazy_fn(int(i))

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def extend(self, payloads):
        self._items.extend(payloads)

    def set_lazy(self, n, fn):
        if len(self):
            raise RetrievalError("lazy payloads can only be attached to an empty store")
        self._lazy_n, self._lazy_fn = int(n), fn

    def drop(self, ids):
        """Forget the payloads of physically dropped rows (ids stay allocated: row ids are never reused)."""
        for i in ids:
            if i >= self._lazy_n:
                self._items[int(i) - self._lazy_n] = None

    def clear(self):
        self._lazy_n, self._lazy_fn, self._items = 0, None, []


@_register
class B200Retriever(BaseRetriever):
    """Exact dense / sparse / hybrid-RRF retrieval on one or several B200 shards owned by this process."""

    def __init__(self, config: RetrievalConfig, embedding_dim: int = 1024, *, device: int = 0,
                 devices: list[int] | str | None = None, vocab: int = 250_002, docs_per_block: int = 0,
                 rrf_k: int = 2, row_base: int = 0, compact_dead_fraction: float = 0.25,
                 device_add_rows: int = 1024, compressed_scan: bool | None = None):
        self.config = config
        self.embedding_dim = embedding_dim
        self._vocab, self._R, self._rrf_k, self._row_base = vocab, docs_per_block, rrf_k, row_base
        if devices is None:
            devices = os.environ.get("B200RAG_DEVICES") or None
        self._devices_spec = devices if devices is not None else [device]
        self._compact_dead_fraction = float(compact_dead_fraction)
        self._device_add_rows = int(device_add_rows)
        # opt-in 8-bit candidate scan (b200rag_set_compression): same results, half the bytes per search, +50 % HBM
        if compressed_scan is None:
            compressed_scan = os.environ.get("B200RAG_COMPRESSED_SCAN", "0") not in ("", "0")
        self._compressed_scan = bool(compressed_scan)
        self._pack_dense = None                 # struct.Struct("<dim>f"), built on first use
        self._shards: list | None = None
        self._group = None
        self._existing_collections: set[str] = set()
        self._hybrid_collections: set[str] = set()
        self._reset_rows()
        self._coll_ids: dict[str, int] = {}
        self._coll_rows: dict[str, int] = {}    # live rows per collection
        self._coll_version: dict[str, int] = {}
        logger.info(f"B200Retriever initialized: collection={config.collection_name}, "
                    f"search_type={config.search_type}")

    def _reset_rows(self):
        """Empty host-side row bookkeeping (fresh retriever, last collection deleted, load)."""
        self._payloads = _PayloadStore()
        self._coll_of = _Grow(np.int32)         # collection id per global row id
        self._alive_of = _Grow(bool)            # tombstones
        self._shard_rows: list[_Grow] = []      # per shard: global ids of its local rows, in local order
        self._stored = 0                        # rows physically held by the shards (live or tombstoned)
        self._masks: OrderedDict = OrderedDict()  # (collection, filter key) -> (mask_id, version)
        self._meta_index: dict = {}               # metadata key -> {"upto": ids indexed, "map": value -> [ids], "slow": [ids]}
        self._next_mask = 0
        self._layout_epoch = 0                  # bumped when local rows move (compaction): every cached mask is void

    # views used by the tests and by callers that want to look at the row state
    @property
    def _row_coll(self):
        return self._coll_of.view

    @property
    def _alive(self):
        return self._alive_of.view

    # ------------------------------------------------------------------ engine handles (qdrant.py:35-54)
    def _resolve_devices(self) -> list[int]:
        spec = self._devices_spec
        if isinstance(spec, str):
            if spec.strip().lower() == "all":
                n = _ffi.device_count()
                if n < 1:
                    raise RetrievalError("B200RAG_DEVICES=all but no sm_100 device is visible")
                return list(range(n))
            return [int(x) for x in spec.replace(" ", "").split(",") if x != ""]
        return [int(x) for x in spec]

    def _get_shards(self) -> list:
        if self._shards is not None:
            return self._shards
        try:
            devs = self._resolve_devices()
            if not devs:
                raise RetrievalError("no device given")
            shards = [_ffi.Shard(dim=self.embedding_dim, vocab=self._vocab, device=d, row_base=self._row_base,
                                 docs_per_block=self._R) for d in devs]
            self._set_shards(shards)
            return self._shards
        except Exception as e:
            # same wording as the reference so the API layer's "connect" -> 503 mapping keeps working (api/v1/query.py:151-161)
            raise RetrievalError(f"Failed to connect to the B200 retrieval engine: {e}")

    def _set_shards(self, shards, group=None):
        self._shards = list(shards)
        if getattr(self, "_compressed_scan", False):
            for sh in self._shards:
                if hasattr(sh, "set_compression"):
                    try:
                        sh.set_compression(True)
                    except Exception as e:      # (e.g. an embedding width the 8-bit scan does not cover: same results without it)
                        logger.warning(f"compressed candidate scan not available, continuing with the bf16 scan: {e}")
                        break
        while len(self._shard_rows) < len(self._shards):
            self._shard_rows.append(_Grow(np.int64))
        if group is None and len(self._shards) > 1:
            group = _ffi.ShardGroup(self._shards)
        self._group = group

    # single-shard handle (kept for callers/tests that inject a shard or a test double)
    @property
    def _shard(self):
        return self._shards[0] if self._shards else None

    @_shard.setter
    def _shard(self, shard):
        self._set_shards([shard])

    def _get_shard(self):
        return self._get_shards()[0]

    @property
    def n_shards(self) -> int:
        return len(self._get_shards())

    def close(self):
        if self._group is not None and hasattr(self._group, "close"):
            self._group.close()
        self._group = None
        if self._shards is not None:
            for s in self._shards:
                s.close()
            self._shards = None

    def _resolve_collection(self, collection_name: str | None) -> str:
        return collection_name or self.config.collection_name

    def _ensure_collection(self, collection_name: str | None = None, hybrid: bool = False) -> str:
        """qdrant.py:59-132: unknown names are created (dense-only unless ``hybrid``), known ones keep their schema."""
        resolved = self._resolve_collection(collection_name)
        if resolved in self._existing_collections:
            if hybrid and resolved not in self._hybrid_collections:
                logger.warning(f"Collection {resolved} exists but is not hybrid-enabled. "
                               "Re-index required for hybrid search.")
            return resolved
        logger.info(f"Creating {'hybrid' if hybrid else 'dense'} collection: {resolved}")
        if resolved not in self._coll_ids:
            self._coll_ids[resolved] = len(self._coll_ids)
        self._coll_rows[resolved] = 0
        self._coll_version[resolved] = self._coll_version.get(resolved, 0) + 1
        if hybrid:
            self._hybrid_collections.add(resolved)
        self._existing_collections.add(resolved)
        return resolved

    def is_hybrid_collection(self, collection_name: str | None = None) -> bool:
        resolved = self._resolve_collection(collection_name)
        self._ensure_collection(resolved)
        return resolved in self._hybrid_collections

    # ------------------------------------------------------------------ add (qdrant.py:140-225)
    def _prepare_rows(self, chunks, embeddings, is_hybrid):
        """Host staging of an add(): payloads, raw fp32 rows, doc-major CSR of the sparse parts."""
        dense = np.asarray([e.dense for e in embeddings], dtype=np.float32)
        if dense.ndim != 2 or dense.shape[1] != self.embedding_dim:
            raise RetrievalError(f"dense vectors must have dimension {self.embedding_dim}")
        indptr = np.zeros(len(chunks) + 1, dtype=np.int64)
        tt, ww = [], []
        for i, emb in enumerate(embeddings):
            n = 0
            if is_hybrid and emb.sparse is not None:
                t, w = _sorted_sparse(emb.sparse, self._vocab)
                tt.append(t)
                ww.append(w)
                n = len(t)
            indptr[i + 1] = indptr[i] + n
        terms = np.concatenate(tt) if tt else np.zeros(0, np.uint32)
        weights = np.concatenate(ww) if ww else np.zeros(0, np.float32)
        payloads = [{"text": c.text, "start": c.start, "end": c.end, "speaker": c.speaker,
                     "metadata": c.metadata or {}} for c in chunks]
        return dense, indptr, terms, weights, payloads

    def _route(self, n: int) -> list[tuple[int, int, int]]:
        """Which shard takes which contiguous slice of an add() batch: small batches go whole to the emptiest shard,
        bulk batches are cut into one slice per shard (ids stay increasing inside every shard either way)."""
        ns = len(self._shards)
        if ns == 1:
            return [(0, 0, n)]
        loads = [g.n for g in self._shard_rows]
        if n < 4096 * ns:
            return [(int(np.argmin(loads)), 0, n)]
        # water-fill: after the add every shard should hold about the same number of rows
        target = (sum(loads) + n) / ns
        want = np.maximum(0, np.floor(target - np.asarray(loads))).astype(np.int64)
        short = n - int(want.sum())
        order = np.argsort(loads, kind="stable")
        for i in range(abs(short)):
            want[order[i % ns]] += 1 if short > 0 else -1 if want[order[i % ns]] > 0 else 0
        while want.sum() != n:                       # (only when the -1 branch skipped an empty slice)
            j = int(np.argmax(want))
            want[j] += n - int(want.sum())
        parts, lo = [], 0
        for s in range(ns):
            if want[s] > 0:
                parts.append((s, lo, lo + int(want[s])))
                lo += int(want[s])
        return parts

    @timed
    def add(self, chunks: list[AudioChunk], embeddings: list[EmbeddingResult],
            collection_name: str | None = None) -> None:
        if not chunks:
            return
        if len(chunks) != len(embeddings):
            raise RetrievalError(f"Chunks/embeddings mismatch: {len(chunks)} chunks, {len(embeddings)} embeddings")
        has_sparse = any(e.sparse is not None for e in embeddings)
        resolved = self._ensure_collection(collection_name, hybrid=has_sparse)
        try:
            is_hybrid = resolved in self._hybrid_collections
            dense, indptr, terms, weights, payloads = self._prepare_rows(chunks, embeddings, is_hybrid)
            shards = self._get_shards()
            cid = self._coll_ids[resolved]
            n = len(chunks)
            g0 = self._row_base + len(self._payloads)
            ids = np.arange(g0, g0 + n, dtype=np.int64)
            # large batches: the GPU normalises + packs the raw fp32 rows (b200rag_add_f32); small ones use the host
            # routine (bit-equal) and skip the fp32 upload
            on_device = n >= self._device_add_rows
            bits = None if on_device else _ffi.normalize_bf16(dense)
            for si, lo, hi in self._route(n):
                ip = indptr[lo:hi + 1] - indptr[lo]
                tt, ww = terms[indptr[lo]:indptr[hi]], weights[indptr[lo]:indptr[hi]]
                if on_device:
                    shards[si].add_f32(dense[lo:hi], ip, tt, ww, ids=ids[lo:hi])
                else:
                    shards[si].add(bits[lo:hi], ip, tt, ww, ids=ids[lo:hi])
                # bookkeeping per slice, so a failure on a later shard leaves rows and books consistent
                self._payloads.extend(payloads[lo:hi])
                self._coll_of.fill(hi - lo, cid)
                self._alive_of.fill(hi - lo, True)
                self._shard_rows[si].append(ids[lo:hi])
                self._stored += hi - lo
                self._coll_rows[resolved] += hi - lo
                self._coll_version[resolved] += 1
            logger.info(f"Added {len(chunks)} chunks to {resolved} (hybrid={is_hybrid})")
        except RetrievalError as e:
            raise RetrievalError(f"Failed to add chunks to '{resolved}': {e}")
        except Exception as e:
            raise RetrievalError(f"Failed to add chunks to '{resolved}': {e}")

    # ------------------------------------------------------------------ eligibility masks (R4)
    def _eligible(self, resolved: str, filter_metadata: dict | None) -> np.ndarray | None:
        """bool per global row id, or None when every STORED row is eligible (one live collection, no filter)."""
        n = len(self._payloads)
        if not filter_metadata and self._coll_rows.get(resolved, 0) == self._stored:
            return None
        cid = self._coll_ids[resolved]
        elig = (self._row_coll == cid) & self._alive
        for k, v in (filter_metadata or {}).items():
            hit = np.zeros(n, dtype=bool)
            hit[self._meta_rows(k, v) - self._row_base] = True
            elig &= hit
        return elig

    def _meta_rows(self, key, value) -> np.ndarray:
        """Global ids of the rows (of any collection) whose ``metadata[key]`` matches ``value`` under ``_match``.  A
        per-key payload index (value -> rows, list-valued fields indexed by element) is built on the first filter that
        names the key and extended as rows are added -- the host-side analogue of a qdrant payload index, so that a
        new filter costs one lookup instead of a Python pass over every payload."""
        idx = self._meta_index.setdefault(key, {"upto": 0, "map": {}, "slow": []})
        base = self._row_base
        for r in range(idx["upto"], len(self._payloads)):
            p = self._payloads[r]
            meta = p.get("metadata") if p is not None else None
            if not isinstance(meta, dict) or key not in meta:
                continue
            got = meta[key]
            try:
                for item in (got if isinstance(got, (list, tuple)) else (got,)):
                    rows = idx["map"].setdefault(item, [])
                    if not rows or rows[-1] != r + base:
                        rows.append(r + base)
            except TypeError:                     # unhashable stored value: matched the slow way
                idx["slow"].append(r + base)
        idx["upto"] = len(self._payloads)
        try:
            rows = list(idx["map"].get(value, ()))
            slow = idx["slow"]
        except TypeError:                         # unhashable filter value: every row that has the key is a candidate
            rows, slow = [], sorted({r for rr in idx["map"].values() for r in rr} | set(idx["slow"]))
        for r in slow:
            p = self._payloads[r - base]
            if p is not None and _match(p.get("metadata"), key, value):
                rows.append(r)
        return np.asarray(rows, dtype=np.int64)

    def _mask_id(self, resolved: str, filter_metadata: dict | None) -> int:
        """Mask id (the same on every shard) for (collection, filter), uploading it if it is not cached; -1 = all rows.
        Nothing is evicted here: `_execute` trims the cache after the search that may still use the ids."""
        elig_key = (resolved, tuple(sorted((str(k), repr(v)) for k, v in (filter_metadata or {}).items())))
        version = (self._coll_version[resolved], len(self._payloads), self._layout_epoch)
        hit = self._masks.get(elig_key)
        if hit is not None and hit[1] == version:
            self._masks.move_to_end(elig_key)
            return hit[0]
        elig = self._eligible(resolved, filter_metadata)
        if elig is None:
            return -1
        shards = self._get_shards()
        if hit is not None:
            mid = hit[0]
        else:
            mid = self._next_mask
            self._next_mask += 1
        for s, sh in enumerate(shards):
            local = self._shard_rows[s].view - self._row_base          # global ids of the shard's local rows
            sh.mask_set(mid, pack_mask(elig[local]), len(local))
        self._masks[elig_key] = (mid, version)
        self._masks.move_to_end(elig_key)
        return mid

    def _evict_masks(self, pinned: set) -> None:
        """LRU eviction, never of a mask the batch that just ran was using (its ids were already handed to the engine)."""
        shards = self._get_shards()
        limit = max(_MASK_CACHE, len(pinned))
        for key in list(self._masks):
            if len(self._masks) <= limit:
                break
            mid = self._masks[key][0]
            if mid in pinned:
                continue
            del self._masks[key]
            for sh in shards:
                sh.mask_drop(mid)

    # ------------------------------------------------------------------ search (qdrant.py:227-352)
    def _plan_search(self, query_embedding, top_k, collection_name, filter_metadata, search_type) -> dict:
        """Pure host logic: defaults + branch selection exactly as qdrant.py:248-332 (rule R11)."""
        resolved = self._ensure_collection(collection_name)
        top_k = top_k or self.config.top_k
        search_type = search_type or self.config.search_type
        is_hybrid = resolved in self._hybrid_collections
        has_sparse = query_embedding.sparse is not None  # `if query_embedding.sparse` on a dataclass instance
        threshold = None
        if search_type == "hybrid" and is_hybrid and has_sparse:
            mode, leg_limit = "hybrid", top_k * 2
        elif search_type == "sparse" and is_hybrid and has_sparse:
            mode, leg_limit = "sparse", top_k
        else:
            mode, leg_limit = "dense", top_k
            if not is_hybrid and self.config.score_threshold > 0:
                threshold = float(self.config.score_threshold)
        return {"collection": resolved, "mode": mode, "top_k": int(top_k), "leg_limit": int(leg_limit),
                "score_threshold": threshold, "filter": dict(filter_metadata) if filter_metadata else None}

    def _materialise(self, ids, scores, count, resolved) -> list[RetrievalResult]:
        n = int(count)
        if n == 0:
            return []
        out = []
        payloads, base = self._payloads, self._row_base
        # (one tolist() per array instead of a numpy scalar + int()/float() per hit)
        for rid, sc in zip(np.asarray(ids[:n]).tolist(), np.asarray(scores[:n], dtype=np.float64).tolist()):
            payload = payloads[rid - base]
            get = payload.get
            md = get("metadata")
            out.append(RetrievalResult(chunk=AudioChunk(text=get("text", ""), start=get("start", 0.0), end=get("end", 0.0),
                                                        speaker=get("speaker"),
                                                        metadata=dict(md) if md is not None else None),
                                       score=sc, source=resolved))
        return out

    def _query_arrays(self, embeddings, mode):
        if len(embeddings) == 1 and isinstance(embeddings[0].dense, list):
            # the single-query call of the reference (a Python list of floats): struct.pack converts it 2.6x as fast as
            # numpy does (16 vs 41 us for 1024 floats; array('f'): 27 us) with the same double -> float rounding
            try:
                pk = self._pack_dense
                if pk is None:
                    pk = self._pack_dense = struct.Struct(f"{self.embedding_dim}f")
                dense = np.frombuffer(pk.pack(*embeddings[0].dense), dtype=np.float32).reshape(1, -1)
            except (struct.error, TypeError, OverflowError):
                dense = np.asarray([embeddings[0].dense], dtype=np.float32)
        else:
            dense = np.asarray([e.dense for e in embeddings], dtype=np.float32)
        if dense.ndim != 2 or dense.shape[1] != self.embedding_dim:
            raise RetrievalError(f"query vectors must have dimension {self.embedding_dim}")
        q_bits = _ffi.normalize_bf16(dense)
        if mode == "dense":
            return q_bits, None, None, None
        if len(embeddings) == 1:                    # the reference's single-query call: nothing to concatenate
            t, w = _sorted_sparse(embeddings[0].sparse, self._vocab)
            return q_bits, np.array([0, len(t)], dtype=np.int64), t, w
        indptr = np.zeros(len(embeddings) + 1, dtype=np.int64)
        tt, ww = [], []
        for i, e in enumerate(embeddings):
            t, w = _sorted_sparse(e.sparse, self._vocab)
            tt.append(t)
            ww.append(w)
            indptr[i + 1] = indptr[i] + len(t)
        return (q_bits, indptr, np.concatenate(tt) if tt else np.zeros(0, np.uint32),
                np.concatenate(ww) if ww else np.zeros(0, np.float32))

    @timed
    def search(self, query_embedding: EmbeddingResult, top_k: int | None = None,
               collection_name: str | None = None, filter_metadata: dict | None = None,
               search_type: str | None = None) -> list[RetrievalResult]:
        plan = self._plan_search(query_embedding, top_k, collection_name, filter_metadata, search_type)
        resolved = plan["collection"]
        try:
            return self._execute([query_embedding], [plan])[0]
        except Exception as e:
            raise RetrievalError(f"Search failed in '{resolved}': {e}")

    def search_batch(self, query_embeddings: list[EmbeddingResult], top_k: int | None = None,
                     collection_name: str | list[str] | None = None, filter_metadata: dict | None = None,
                     search_type: str | None = None) -> list[list[RetrievalResult]]:
        """Additive: many queries, one pass over the corpus.  ``collection_name`` may be one name or one per query."""
        names = collection_name if isinstance(collection_name, (list, tuple)) else [collection_name] * len(query_embeddings)
        if len(names) != len(query_embeddings):
            raise RetrievalError("search_batch: one collection name per query expected")
        plans = [self._plan_search(q, top_k, n, filter_metadata, search_type) for q, n in zip(query_embeddings, names)]
        try:
            return self._execute(query_embeddings, plans)
        except Exception as e:
            raise RetrievalError(f"Search failed: {e}")

    def search_batch_arrays(self, query_embeddings: list[EmbeddingResult], top_k: int | None = None,
                            collection_name: str | list[str] | None = None, filter_metadata: dict | None = None,
                            search_type: str | None = None) -> dict:
        """Additive (the reranker hand-off, reranking/bge.py:86-147, pipeline/query.py:145-160):
        the same search as ``search_batch`` without building ``RetrievalResult``/``AudioChunk`` objects per hit.
        Returns ``{"ids": int64 [B, k] (-1 padded), "scores": float64 [B, k], "counts": int32 [B], "texts": list of B
        lists of chunk texts, "sources": list of B collection names}`` -- what a cross-encoder needs to build its
        (query, passage) pairs in one go; ``materialise(b, j)`` turns any hit into the usual ``RetrievalResult``."""
        names = collection_name if isinstance(collection_name, (list, tuple)) else [collection_name] * len(query_embeddings)
        if len(names) != len(query_embeddings):
            raise RetrievalError("search_batch_arrays: one collection name per query expected")
        plans = [self._plan_search(q, top_k, n, filter_metadata, search_type) for q, n in zip(query_embeddings, names)]
        k = plans[0]["top_k"] if plans else 0
        B = len(plans)
        ids = np.full((B, k), -1, dtype=np.int64)
        scores = np.zeros((B, k), dtype=np.float64)
        counts = np.zeros(B, dtype=np.int32)
        try:
            raw = self._execute(query_embeddings, plans, raw=True)
        except Exception as e:
            raise RetrievalError(f"Search failed: {e}")
        texts = []
        for b, (i, s, c) in enumerate(raw):
            ids[b, :c], scores[b, :c], counts[b] = i[:c], s[:c], c
            texts.append([self._payloads[int(r) - self._row_base].get("text", "") for r in i[:c]])
        out = {"ids": ids, "scores": scores, "counts": counts, "texts": texts,
               "sources": [p["collection"] for p in plans]}
        out["materialise"] = lambda b, j: self._materialise(ids[b, j:j + 1], scores[b, j:j + 1], 1, out["sources"][b])[0]
        return out

    def _search_engine(self, mode, k, q_bits, indptr, terms, weights, mask_ids, thr):
        """One engine call for a homogeneous batch: the single shard, or the group over all shards."""
        target = self._shards[0] if len(self._shards) == 1 else self._group
        return target.search(mode, k, q_bits, indptr, terms, weights, mask_ids=mask_ids, score_threshold=thr,
                             rrf_k=self._rrf_k)

    def _execute(self, embeddings, plans, raw: bool = False) -> list:
        results: list = [None] * len(plans)
        groups: dict = {}
        for i, p in enumerate(plans):
            if self._coll_rows.get(p["collection"], 0) == 0:
                # the engine is still required to exist: a missing GPU/library must not look like "no results"
                self._get_shards()
                results[i] = (np.zeros(0, np.int64), np.zeros(0, np.float64), 0) if raw else []
                continue
            groups.setdefault((p["mode"], p["top_k"], p["score_threshold"]), []).append(i)
        for (mode, k, thr), idxs in groups.items():
            self._get_shards()
            q_bits, indptr, terms, weights = self._query_arrays([embeddings[i] for i in idxs], mode)
            mask_ids = np.asarray([self._mask_id(plans[i]["collection"], plans[i]["filter"]) for i in idxs], np.int32)
            try:
                ids, scores, counts = self._search_engine(mode, k, q_bits, indptr, terms, weights,
                                                          mask_ids if (mask_ids >= 0).any() else None, thr)
            finally:
                if len(self._masks) > _MASK_CACHE:
                    self._evict_masks({int(m) for m in mask_ids if m >= 0})
            for j, i in enumerate(idxs):
                results[i] = (ids[j], scores[j], int(counts[j])) if raw else \
                    self._materialise(ids[j], scores[j], counts[j], plans[i]["collection"])
        return results

    # ------------------------------------------------------------------ admin (qdrant.py:354-381)
    def delete_collection(self, collection_name: str | None = None) -> None:
        """qdrant.py:354-363 drops the collection's storage.  Here the rows are tombstoned through the masks at once and
        PHYSICALLY dropped (dense rows, forward and inverted index, on the device) as soon as a shard's dead fraction
        exceeds ``compact_dead_fraction``; when nothing is live anywhere everything is dropped."""
        resolved = self._resolve_collection(collection_name)
        try:
            if resolved in self._existing_collections:
                cid = self._coll_ids[resolved]
                self._alive[self._row_coll == cid] = False
                self._coll_rows[resolved] = 0
                self._coll_version[resolved] += 1
                if not self._alive.any():
                    # nothing live anywhere: drop the rows for real
                    if self._shards is not None:
                        for s in self._shards:
                            s.clear()
                    n_shards = len(self._shard_rows)
                    self._reset_rows()
                    self._shard_rows = [_Grow(np.int64) for _ in range(n_shards)]
                else:
                    self._compact()
            self._existing_collections.discard(resolved)
            self._hybrid_collections.discard(resolved)
            logger.info(f"Deleted collection: {resolved}")
        except Exception as e:
            raise RetrievalError(f"Failed to delete collection '{resolved}': {e}")

    def _compact(self, force: bool = False) -> int:
        """Physically drop tombstoned rows on every shard whose dead fraction exceeds the threshold; returns the
        number of rows dropped.  Global ids do not change (the shards keep their id maps), only local positions do,
        so every cached mask is invalidated."""
        if self._shards is None:
            return 0
        dropped = 0
        base = self._row_base
        for s, sh in enumerate(self._shards):
            rows = self._shard_rows[s].view
            if len(rows) == 0:
                continue
            keep = self._alive[rows - base]
            dead = int(len(rows) - keep.sum())
            if dead == 0 or (not force and dead <= self._compact_dead_fraction * len(rows)):
                continue
            sh.compact(pack_mask(keep), len(rows))
            self._payloads.drop(rows[~keep] - base)
            self._shard_rows[s].assign(rows[keep])
            self._stored -= dead
            dropped += dead
        if dropped:
            self._layout_epoch += 1
            self._masks.clear()           # (the engine dropped its masks with the rows)
            logger.info(f"Compacted {dropped} deleted rows")
        return dropped

    # ------------------------------------------------------------------ bulk attach (additive; bench / offline builds)
    def attach_prebuilt(self, shards, collection_name: str | None = None, *, payload_fn=None, hybrid: bool = True,
                        group=None) -> None:
        """Adopt shards that were filled OUTSIDE the plugin (bulk / synthetic ingest straight into device memory):
        their rows become ONE collection of this (empty) retriever.  Every shard must hold the contiguous id range
        [row_base_s, row_base_s + count_s), the ranges back to back from this retriever's ``row_base``.
        ``payload_fn(global_id) -> payload dict`` supplies payloads on demand (no per-row Python objects are kept)."""
        if len(self._payloads) or self._shards is not None:
            raise RetrievalError("attach_prebuilt() needs an empty retriever")
        resolved = self._ensure_collection(collection_name, hybrid=hybrid)
        self._set_shards(shards, group)
        lo = self._row_base
        for s, sh in enumerate(self._shards):
            n = int(sh.count)
            if int(sh.row_base) != lo:
                raise RetrievalError("attach_prebuilt: shard id ranges must be contiguous from row_base")
            self._shard_rows[s].assign(np.arange(lo, lo + n, dtype=np.int64))
            lo += n
        total = lo - self._row_base
        fn = payload_fn or (lambda i: {"text": f"row {i}", "start": 0.0, "end": 0.0, "speaker": None, "metadata": {}})
        base = self._row_base
        self._payloads.set_lazy(total, lambda i: fn(i + base))
        self._coll_of.fill(total, self._coll_ids[resolved])
        self._alive_of.fill(total, True)
        self._stored = total
        self._coll_rows[resolved] = total
        self._coll_version[resolved] += 1

    # ------------------------------------------------------------------ persistence (additive)
    # The reference keeps its index in Qdrant's volume (docker-compose.yml:36-37).  Here a retriever is a directory:
    #   shard-<i>.bin        rows + forward sparse index + global row ids of shard i (b200rag_save; the inverted index
    #                        is rebuilt on load)
    #   payloads-<j>.json    payloads of global ids [j * 65536, (j + 1) * 65536) as ONE JSON array (null = dropped row)
    #   rows.npz             per global id: collection id, tombstone; per shard: the global ids of its local rows
    #   manifest.json        format, geometry, collections (written LAST: a directory without it is not a snapshot)
    # Every file is written under a temporary name and renamed into place.
    def save(self, directory: str) -> None:
        try:
            os.makedirs(directory, exist_ok=True)
            shards = self._get_shards()

            def put(name, writer):
                tmp = os.path.join(directory, f".{name}.tmp-{os.getpid()}")
                writer(tmp)
                os.replace(tmp, os.path.join(directory, name))

            for i, sh in enumerate(shards):
                put(f"shard-{i}.bin", sh.save)
            n = len(self._payloads)
            n_chunks = (n + _PAYLOAD_CHUNK - 1) // _PAYLOAD_CHUNK
            for j in range(n_chunks):
                block = [self._payloads[i] for i in range(j * _PAYLOAD_CHUNK, min(n, (j + 1) * _PAYLOAD_CHUNK))]

                def w(tmp, block=block):
                    with open(tmp, "w", encoding="utf-8") as f:
                        json.dump(block, f, ensure_ascii=False)
                put(f"payloads-{j}.json", w)

            def wrows(tmp):
                with open(tmp, "wb") as f:
                    np.savez(f, row_collection=self._row_coll, alive=self._alive,
                             **{f"shard_rows_{i}": g.view for i, g in enumerate(self._shard_rows)})
            put("rows.npz", wrows)
            manifest = {
                "format": "b200rag-retriever-2", "embedding_dim": self.embedding_dim, "vocab": self._vocab,
                "row_base": self._row_base, "rows": n, "shards": len(shards), "payload_chunks": n_chunks,
                "payload_chunk_rows": _PAYLOAD_CHUNK,
                "collections": {name: {"id": cid, "hybrid": name in self._hybrid_collections,
                                       "exists": name in self._existing_collections,
                                       "live_rows": self._coll_rows.get(name, 0)}
                                for name, cid in self._coll_ids.items()},
            }

            def wman(tmp):
                with open(tmp, "w", encoding="utf-8") as f:
                    json.dump(manifest, f)
            put("manifest.json", wman)
        except Exception as e:
            raise RetrievalError(f"Failed to save retriever to '{directory}': {e}")

    def load(self, directory: str) -> None:
        """Restore a saved retriever into this one, which must hold no rows; its collection registry is REPLACED by the
        saved one (names registered earlier, e.g. by a ``count()`` health probe, do not survive).  embedding_dim,
        vocab and the number of shards must match the saved ones."""
        try:
            if len(self._payloads) or self._stored:
                raise RetrievalError("load() needs a retriever without rows")
            with open(os.path.join(directory, "manifest.json"), encoding="utf-8") as f:
                m = json.load(f)
            if m.get("format") != "b200rag-retriever-2" or m["embedding_dim"] != self.embedding_dim or \
                    m["vocab"] != self._vocab:
                raise RetrievalError("manifest does not match this retriever (format, embedding_dim or vocab)")
            shards = self._get_shards()
            if m["shards"] != len(shards):
                raise RetrievalError(f"snapshot has {m['shards']} shards, this retriever has {len(shards)}")
            payloads: list = []
            for j in range(m["payload_chunks"]):
                with open(os.path.join(directory, f"payloads-{j}.json"), encoding="utf-8") as f:
                    payloads.extend(json.load(f))
            z = np.load(os.path.join(directory, "rows.npz"), allow_pickle=False)
            row_coll, alive = z["row_collection"], z["alive"]
            if len(payloads) != m["rows"] or len(row_coll) != m["rows"] or len(alive) != m["rows"]:
                raise RetrievalError("payload files / rows.npz / manifest.json row counts disagree")
            shard_rows = [z[f"shard_rows_{i}"] for i in range(len(shards))]
            try:
                for i, sh in enumerate(shards):
                    sh.load(os.path.join(directory, f"shard-{i}.bin"))
                    if sh.count != len(shard_rows[i]):
                        raise RetrievalError(f"shard-{i}.bin holds a different number of rows than rows.npz")
            except Exception:
                for sh in shards:
                    sh.clear()
                raise
            # everything was read and checked: replace the registry and the row books in one go
            self._reset_rows()
            self._row_base = m["row_base"]
            self._payloads.extend(payloads)
            self._coll_of.assign(row_coll)
            self._alive_of.assign(alive)
            self._shard_rows = []
            for rows in shard_rows:
                g = _Grow(np.int64)
                g.assign(rows)
                self._shard_rows.append(g)
            self._stored = int(sum(len(r) for r in shard_rows))
            self._coll_ids, self._coll_rows = {}, {}
            self._existing_collections, self._hybrid_collections = set(), set()
            for name, c in m["collections"].items():
                self._coll_ids[name] = int(c["id"])
                self._coll_rows[name] = int(c["live_rows"])
                self._coll_version[name] = self._coll_version.get(name, 0) + 1
                if c["exists"]:
                    self._existing_collections.add(name)
                if c["hybrid"]:
                    self._hybrid_collections.add(name)
        except RetrievalError as e:
            raise RetrievalError(f"Failed to load retriever from '{directory}': {e}")
        except Exception as e:
            raise RetrievalError(f"Failed to load retriever from '{directory}': {e}")

    def count(self, collection_name: str | None = None) -> int:
        resolved = self._ensure_collection(collection_name)
        return int(self._coll_rows.get(resolved, 0))

    def collection_exists(self, collection_name: str | None = None) -> bool:
        return self._resolve_collection(collection_name) in self._existing_collections
