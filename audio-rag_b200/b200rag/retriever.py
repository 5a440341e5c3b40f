"""B200Retriever: the reference's retrieval plugin interface on top of the CUDA shard engine.

Mirror of ``QdrantRetriever`` (/root/reference/src/audio_rag/retrieval/qdrant.py:14-381): same constructor,
same ``add / search / count / collection_exists / delete_collection / is_hybrid_collection`` surface, same
defaults (``collection_name or config.collection_name`` :56-57, ``top_k or config.top_k`` :249,
``search_type or config.search_type`` :250), same branch/fallback rules (:272, :299, :313), the same
``limit = 2 * top_k`` per hybrid leg (:287, :292) and the same error convention (every failure surfaces as
``RetrievalError``).  What the reference delegates to qdrant-client (scoring, candidate selection, RRF) runs
in libb200rag.so; what stays on the host is what is host work in the reference too: the payload store,
the collection registry and result materialisation (:334-346).

All collections share one global row space (row id = insertion order); a collection, its
tombstones and an optional ``filter_metadata`` are expressed as one eligibility bitmask per search
(north star: "the collection filter applied as a bitmask"), applied in BOTH legs.

Additive API (no reference counterpart): ``search_batch``.
"""
from __future__ import annotations

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


def _sorted_sparse(sv, vocab):
    """indices/values lists -> (uint32 ascending, float32); duplicates and out-of-range indices are rejected (R3)."""
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


@_register
class B200Retriever(BaseRetriever):
    """Exact dense / sparse / hybrid-RRF retrieval on one B200 shard (or one rank's shard of a sharded corpus)."""

    def __init__(self, config: RetrievalConfig, embedding_dim: int = 1024, *, device: int = 0,
                 vocab: int = 250_002, docs_per_block: int = 0, rrf_k: int = 2, row_base: int = 0):
        self.config = config
        self.embedding_dim = embedding_dim
        self._device, self._vocab, self._R, self._rrf_k, self._row_base = device, vocab, docs_per_block, rrf_k, row_base
        self._shard: _ffi.Shard | None = None
        self._existing_collections: set[str] = set()
        self._hybrid_collections: set[str] = set()
        # host-side row bookkeeping
        self._payloads: list[dict] = []
        self._row_coll = np.zeros(0, dtype=np.int32)   # collection id per row   } views of length len(_payloads) into
        self._alive = np.zeros(0, dtype=bool)          # tombstones              } buffers that grow by doubling
        self._row_cap = (np.zeros(1024, dtype=np.int32), np.zeros(1024, dtype=bool))
        self._coll_ids: dict[str, int] = {}
        self._coll_rows: dict[str, int] = {}    # live rows per collection
        self._coll_version: dict[str, int] = {}
        self._masks: OrderedDict = OrderedDict()  # (collection, filter key) -> (mask_id, version)
        self._meta_index: dict = {}               # metadata key -> {"upto": rows indexed, "map": value -> [rows], "slow": [rows]}
        self._next_mask = 0
        logger.info(f"B200Retriever initialized: collection={config.collection_name}, "
                    f"search_type={config.search_type}")

    # ------------------------------------------------------------------ engine handle (qdrant.py:35-54)
    def _get_shard(self) -> _ffi.Shard:
        if self._shard is not None:
            return self._shard
        try:
            self._shard = _ffi.Shard(dim=self.embedding_dim, vocab=self._vocab, device=self._device,
                                     row_base=self._row_base, docs_per_block=self._R)
            return self._shard
        except Exception as e:
            # same wording as the reference so the API layer's "connect" -> 503 mapping keeps working (api/v1/query.py:151-161)
            raise RetrievalError(f"Failed to connect to the B200 retrieval engine: {e}")

    def close(self):
        if self._shard is not None:
            self._shard.close()
            self._shard = None

    def _set_rows(self, coll, alive) -> None:
        """Replace the per-row bookkeeping (load / clear)."""
        n = len(coll)
        cap = max(1024, 1 << int(n - 1).bit_length()) if n else 1024
        self._row_cap = (np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=bool))
        self._row_cap[0][:n], self._row_cap[1][:n] = coll, alive
        self._row_coll, self._alive = self._row_cap[0][:n], self._row_cap[1][:n]

    def _append_rows(self, cid: int, n_new: int) -> None:
        """n_new live rows of collection `cid` at the end of the row space (amortised O(1) per row)."""
        n = len(self._row_coll)
        if n + n_new > len(self._row_cap[0]):
            cap = len(self._row_cap[0])
            while cap < n + n_new:
                cap *= 2
            grown = (np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=bool))
            grown[0][:n], grown[1][:n] = self._row_coll, self._alive
            self._row_cap = grown
        self._row_cap[0][n:n + n_new] = cid
        self._row_cap[1][n:n + n_new] = True
        self._row_coll, self._alive = self._row_cap[0][:n + n_new], self._row_cap[1][:n + n_new]

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
        """Host staging of an add(): payloads, unit bf16 rows, doc-major CSR of the sparse parts."""
        dense = np.asarray([e.dense for e in embeddings], dtype=np.float32)
        if dense.ndim != 2 or dense.shape[1] != self.embedding_dim:
            raise RetrievalError(f"dense vectors must have dimension {self.embedding_dim}")
        bits = _ffi.normalize_bf16(dense)
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
        return bits, indptr, terms, weights, payloads

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
            bits, indptr, terms, weights, payloads = self._prepare_rows(chunks, embeddings, is_hybrid)
            self._get_shard().add(bits, indptr, terms, weights)
            cid = self._coll_ids[resolved]
            self._payloads.extend(payloads)
            self._append_rows(cid, len(chunks))
            self._coll_rows[resolved] += len(chunks)
            self._coll_version[resolved] += 1
            logger.info(f"Added {len(chunks)} chunks to {resolved} (hybrid={is_hybrid})")
        except RetrievalError as e:
            raise RetrievalError(f"Failed to add chunks to '{resolved}': {e}")
        except Exception as e:
            raise RetrievalError(f"Failed to add chunks to '{resolved}': {e}")

    # ------------------------------------------------------------------ eligibility masks (R4)
    def _eligible(self, resolved: str, filter_metadata: dict | None) -> np.ndarray | None:
        """bool[n_rows] or None when every stored row is eligible (single live collection, no filter)."""
        n = len(self._payloads)
        if not filter_metadata and self._coll_rows.get(resolved, 0) == n:
            return None
        cid = self._coll_ids[resolved]
        elig = (self._row_coll == cid) & self._alive
        for k, v in (filter_metadata or {}).items():
            hit = np.zeros(n, dtype=bool)
            hit[self._meta_rows(k, v)] = True
            elig &= hit
        return elig

    def _meta_rows(self, key, value) -> np.ndarray:
        """Rows (of any collection) whose ``metadata[key]`` matches ``value`` under ``_match``.  A per-key payload
        index (value -> rows, list-valued fields indexed by element) is built on the first filter that names the key
        and extended as rows are added -- the host-side analogue of a qdrant payload index, so that a new filter costs
        one lookup instead of a Python pass over every payload."""
        idx = self._meta_index.setdefault(key, {"upto": 0, "map": {}, "slow": []})
        for r in range(idx["upto"], len(self._payloads)):
            meta = self._payloads[r].get("metadata")
            if not isinstance(meta, dict) or key not in meta:
                continue
            got = meta[key]
            try:
                for item in (got if isinstance(got, (list, tuple)) else (got,)):
                    rows = idx["map"].setdefault(item, [])
                    if not rows or rows[-1] != r:
                        rows.append(r)
            except TypeError:                     # unhashable stored value: matched the slow way
                idx["slow"].append(r)
        idx["upto"] = len(self._payloads)
        try:
            rows = list(idx["map"].get(value, ()))
            slow = idx["slow"]
        except TypeError:                         # unhashable filter value: every row that has the key is a candidate
            rows, slow = [], sorted({r for rr in idx["map"].values() for r in rr} | set(idx["slow"]))
        rows += [r for r in slow if _match(self._payloads[r].get("metadata"), key, value)]
        return np.asarray(rows, dtype=np.int64)

    def _mask_id(self, resolved: str, filter_metadata: dict | None) -> int:
        elig_key = (resolved, tuple(sorted((str(k), repr(v)) for k, v in (filter_metadata or {}).items())))
        version = (self._coll_version[resolved], len(self._payloads))
        hit = self._masks.get(elig_key)
        if hit is not None and hit[1] == version:
            self._masks.move_to_end(elig_key)
            return hit[0]
        elig = self._eligible(resolved, filter_metadata)
        if elig is None:
            return -1
        shard = self._get_shard()
        if hit is not None:
            mid = hit[0]
        else:
            mid = self._next_mask
            self._next_mask += 1
        shard.mask_set(mid, pack_mask(elig), len(elig))
        self._masks[elig_key] = (mid, version)
        self._masks.move_to_end(elig_key)
        while len(self._masks) > _MASK_CACHE:
            _, (old, _v) = self._masks.popitem(last=False)
            shard.mask_drop(old)
        return mid

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
        out = []
        for j in range(int(count)):
            payload = self._payloads[int(ids[j]) - self._row_base]
            chunk = AudioChunk(text=payload.get("text", ""), start=payload.get("start", 0.0),
                               end=payload.get("end", 0.0), speaker=payload.get("speaker"),
                               metadata=dict(payload["metadata"]) if payload.get("metadata") is not None else None)
            out.append(RetrievalResult(chunk=chunk, score=float(scores[j]), source=resolved))
        return out

    def _query_arrays(self, embeddings, mode):
        dense = np.asarray([e.dense for e in embeddings], dtype=np.float32)
        if dense.ndim != 2 or dense.shape[1] != self.embedding_dim:
            raise RetrievalError(f"query vectors must have dimension {self.embedding_dim}")
        q_bits = _ffi.normalize_bf16(dense)
        if mode == "dense":
            return q_bits, None, None, None
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

    def _execute(self, embeddings, plans, raw: bool = False) -> list:
        results: list = [None] * len(plans)
        groups: dict = {}
        for i, p in enumerate(plans):
            if self._coll_rows.get(p["collection"], 0) == 0:
                # the engine is still required to exist: a missing GPU/library must not look like "no results"
                self._get_shard()
                results[i] = (np.zeros(0, np.int64), np.zeros(0, np.float64), 0) if raw else []
                continue
            groups.setdefault((p["mode"], p["top_k"], p["score_threshold"]), []).append(i)
        for (mode, k, thr), idxs in groups.items():
            shard = self._get_shard()
            q_bits, indptr, terms, weights = self._query_arrays([embeddings[i] for i in idxs], mode)
            mask_ids = np.asarray([self._mask_id(plans[i]["collection"], plans[i]["filter"]) for i in idxs], np.int32)
            ids, scores, counts = shard.search(mode, k, q_bits, indptr, terms, weights,
                                               mask_ids=mask_ids if (mask_ids >= 0).any() else None,
                                               score_threshold=thr, rrf_k=self._rrf_k)
            for j, i in enumerate(idxs):
                results[i] = (ids[j], scores[j], int(counts[j])) if raw else \
                    self._materialise(ids[j], scores[j], counts[j], plans[i]["collection"])
        return results

    # ------------------------------------------------------------------ admin (qdrant.py:354-381)
    def delete_collection(self, collection_name: str | None = None) -> None:
        resolved = self._resolve_collection(collection_name)
        try:
            if resolved in self._existing_collections:
                cid = self._coll_ids[resolved]
                self._alive[self._row_coll == cid] = False
                self._coll_rows[resolved] = 0
                self._coll_version[resolved] += 1
                if not self._alive.any():
                    # nothing live anywhere: drop the rows for real
                    if self._shard is not None:
                        self._shard.clear()
                    self._payloads.clear()
                    self._set_rows([], [])
                    self._masks.clear()
                    self._meta_index.clear()
            self._existing_collections.discard(resolved)
            self._hybrid_collections.discard(resolved)
            logger.info(f"Deleted collection: {resolved}")
        except Exception as e:
            raise RetrievalError(f"Failed to delete collection '{resolved}': {e}")

    # ------------------------------------------------------------------ persistence (additive)
    # The reference keeps its index in Qdrant's volume (docker-compose.yml:36-37).  Here a retriever is three files:
    #   shard.bin       rows + forward sparse index (b200rag_save; the inverted index is rebuilt on load)
    #   payloads.jsonl  one payload per row, insertion order (row id = line number, rule R1)
    #   manifest.json   collections (ids, hybrid flags, live counts), row -> collection, tombstones, geometry
    def save(self, directory: str) -> None:
        try:
            os.makedirs(directory, exist_ok=True)
            self._get_shard().save(os.path.join(directory, "shard.bin"))
            with open(os.path.join(directory, "payloads.jsonl"), "w", encoding="utf-8") as f:
                for p in self._payloads:
                    f.write(json.dumps(p, ensure_ascii=False) + "\n")
            manifest = {
                "format": "b200rag-retriever-1", "embedding_dim": self.embedding_dim, "vocab": self._vocab,
                "row_base": self._row_base, "rows": len(self._payloads),
                "collections": {name: {"id": cid, "hybrid": name in self._hybrid_collections,
                                       "exists": name in self._existing_collections,
                                       "live_rows": self._coll_rows.get(name, 0)}
                                for name, cid in self._coll_ids.items()},
                "row_collection": self._row_coll.tolist(), "alive": self._alive.astype(int).tolist(),
            }
            with open(os.path.join(directory, "manifest.json"), "w", encoding="utf-8") as f:
                json.dump(manifest, f)
        except Exception as e:
            raise RetrievalError(f"Failed to save retriever to '{directory}': {e}")

    def load(self, directory: str) -> None:
        """Restore a saved retriever into this (empty) one; config/embedding_dim/vocab must match the saved ones."""
        try:
            if self._payloads:
                raise RetrievalError("load() needs an empty retriever")
            with open(os.path.join(directory, "manifest.json"), encoding="utf-8") as f:
                m = json.load(f)
            if m.get("format") != "b200rag-retriever-1" or m["embedding_dim"] != self.embedding_dim or \
                    m["vocab"] != self._vocab:
                raise RetrievalError("manifest does not match this retriever (format, embedding_dim or vocab)")
            with open(os.path.join(directory, "payloads.jsonl"), encoding="utf-8") as f:
                payloads = [json.loads(line) for line in f]
            if len(payloads) != m["rows"] or len(m["row_collection"]) != m["rows"] or len(m["alive"]) != m["rows"]:
                raise RetrievalError("payloads.jsonl / manifest.json row counts disagree")
            shard = self._get_shard()
            shard.load(os.path.join(directory, "shard.bin"))
            if shard.count != m["rows"]:
                shard.clear()
                raise RetrievalError("shard.bin holds a different number of rows than the manifest")
            self._row_base = m["row_base"]
            self._payloads = payloads
            self._set_rows(np.asarray(m["row_collection"], dtype=np.int32), np.asarray(m["alive"], dtype=bool))
            for name, c in m["collections"].items():
                self._coll_ids[name] = int(c["id"])
                self._coll_rows[name] = int(c["live_rows"])
                self._coll_version[name] = self._coll_version.get(name, 0) + 1
                if c["exists"]:
                    self._existing_collections.add(name)
                if c["hybrid"]:
                    self._hybrid_collections.add(name)
            self._masks.clear()
            self._meta_index.clear()
        except RetrievalError as e:
            raise RetrievalError(f"Failed to load retriever from '{directory}': {e}")
        except Exception as e:
            raise RetrievalError(f"Failed to load retriever from '{directory}': {e}")

    def count(self, collection_name: str | None = None) -> int:
        resolved = self._ensure_collection(collection_name)
        return int(self._coll_rows.get(resolved, 0))

    def collection_exists(self, collection_name: str | None = None) -> bool:
        return self._resolve_collection(collection_name) in self._existing_collections
