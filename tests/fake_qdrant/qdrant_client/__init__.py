"""TEST DOUBLE for the third-party ``qdrant_client`` package (which is not installable here).

It implements exactly the client surface ``QdrantRetriever`` calls (reference src/audio_rag/retrieval/qdrant.py:
40-50, 87-124, 197-220, 281-332, 358, 369-370, 378-379) on top of the CPU oracle, so that the reference's OWN
plugin code can be executed end to end and compared, request by request and result by result, with
``B200Retriever``.  Semantics follow SURVEY.md rules R1-R12; where qdrant's behaviour was only recalled, not
verified (root filter on fusion queries; upsert of an existing id), the choice made is stated inline.
Every ``query_points`` request is recorded in ``QdrantClient.requests`` for plan-level comparisons.
"""
from __future__ import annotations

import copy
from types import SimpleNamespace

import numpy as np

from oracle import oracle

from . import models
from .models import Distance  # noqa: F401


class _Collection:
    def __init__(self, name, dim, hybrid, legacy_unnamed):
        self.name, self.dim, self.hybrid, self.legacy_unnamed = name, dim, hybrid, legacy_unnamed
        self.index = oracle.OracleIndex(dim)
        self.ids: list = []
        self.row_of: dict = {}
        self.payloads: list = []
        self.dense_f32: list = []
        self.sparse: list = []


class QdrantClient:
    def __init__(self, location=None, host=None, port=None, **kw):
        self.location, self.host, self.port = location, host, port
        self._c: dict[str, _Collection] = {}
        self.requests: list[dict] = []

    # ---- admin
    def get_collections(self):
        return SimpleNamespace(collections=[SimpleNamespace(name=n) for n in self._c])

    def create_collection(self, collection_name, vectors_config=None, sparse_vectors_config=None, **kw):
        if isinstance(vectors_config, dict):
            vp = vectors_config["dense"]
            legacy = False
        else:
            vp, legacy = vectors_config, True
        assert vp.distance == models.Distance.COSINE
        self._c[collection_name] = _Collection(collection_name, vp.size, bool(sparse_vectors_config), legacy)
        return True

    def get_collection(self, collection_name):
        c = self._c[collection_name]
        params = SimpleNamespace(sparse_vectors={"sparse": models.SparseVectorParams()} if c.hybrid else None)
        return SimpleNamespace(points_count=len(c.ids), config=SimpleNamespace(params=params))

    def delete_collection(self, collection_name):
        self._c.pop(collection_name, None)
        return True

    # ---- upsert
    def upsert(self, collection_name, points, **kw):
        c = self._c[collection_name]
        for p in points:
            vec = p.vector
            dense = vec if not isinstance(vec, dict) else vec.get("dense")
            sparse = vec.get("sparse") if isinstance(vec, dict) else None
            if p.id in c.row_of:
                # Existing id.  The reference upserts every hybrid point twice: once with dense+sparse and then
                # again, in the final batch, with the dense vector only (qdrant.py:186-216).  Real qdrant replaces
                # the point on upsert [3P-RECALL], which would drop the sparse vector; the evident intent
                # ("Update sparse separately") is to keep both, and that is what this double and B200Retriever do.
                r = c.row_of[p.id]
                if dense is not None:
                    c.dense_f32[r] = list(dense)
                if sparse is not None:
                    c.sparse[r] = (list(sparse.indices), list(sparse.values))
                c.payloads[r] = copy.deepcopy(p.payload or {})
            else:
                c.row_of[p.id] = len(c.ids)
                c.ids.append(p.id)
                c.dense_f32.append(list(dense))
                c.sparse.append((list(sparse.indices), list(sparse.values)) if sparse is not None else None)
                c.payloads.append(copy.deepcopy(p.payload or {}))
        # rebuild the oracle index (test sizes are small)
        c.index = oracle.OracleIndex(c.dim)
        if c.ids:
            c.index.add_f32(np.asarray(c.dense_f32, dtype=np.float32), c.sparse)
        return True

    # ---- query
    @staticmethod
    def _eligible(c, flt):
        n = len(c.ids)
        elig = np.ones(n, dtype=bool)
        if flt is not None:
            for cond in flt.must:
                assert cond.key.startswith("metadata.")
                k = cond.key[len("metadata."):]
                for r in range(n):
                    meta = (c.payloads[r] or {}).get("metadata") or {}
                    got = meta.get(k, None) if isinstance(meta, dict) else None
                    ok = (k in meta) and ((cond.match.value in got) if isinstance(got, (list, tuple))
                                          else got == cond.match.value)
                    elig[r] = elig[r] and ok
        return elig

    def _leg(self, c, query, using, limit, elig, score_threshold=None):
        if isinstance(query, models.SparseVector):
            assert using == "sparse" and c.hybrid
            qi, qv = oracle.check_sparse_vector(query.indices, query.values)
            ids, scores = c.index.sparse_leg(qi, qv, elig, limit)
        else:
            assert (using == "dense") == (not c.legacy_unnamed), "named vs unnamed vector mismatch"
            qb = oracle.normalize_bf16(np.asarray(query, dtype=np.float32)[None, :])[0]
            ids, scores = c.index.dense_leg(qb, elig, limit, score_threshold)
        return ids, scores

    def query_points(self, collection_name, query=None, using=None, prefetch=None, limit=10, query_filter=None,
                     score_threshold=None, **kw):
        c = self._c[collection_name]
        self.requests.append({
            "collection": collection_name, "limit": limit, "using": using, "score_threshold": score_threshold,
            "fusion": isinstance(query, models.FusionQuery),
            "prefetch": [(p.using, p.limit) for p in (prefetch or [])],
            "filter": None if query_filter is None else {cnd.key: cnd.match.value for cnd in query_filter.must},
            "query_kind": "fusion" if isinstance(query, models.FusionQuery) else
                          ("sparse" if isinstance(query, models.SparseVector) else "dense"),
        })
        if len(c.ids) == 0:
            return SimpleNamespace(points=[])
        # root filter applied inside every prefetch leg (server query-planner semantics)
        elig = self._eligible(c, query_filter)
        if isinstance(query, models.FusionQuery):
            assert query.fusion == models.Fusion.RRF
            legs = [self._leg(c, p.query, p.using, p.limit, elig)[0] for p in prefetch]
            ids, scores = oracle.rrf_fuse(legs, limit)
        else:
            ids, scores = self._leg(c, query, using, limit, elig, score_threshold)
        pts = [models.ScoredPoint(id=c.ids[int(r)], score=float(s), payload=copy.deepcopy(c.payloads[int(r)]))
               for r, s in zip(ids, scores)]
        for p, r in zip(pts, ids):
            p.row = int(r)          # test-only: insertion-order row id (R1) for id-level comparisons
        return SimpleNamespace(points=pts)
