"""Oracle tier T0: the CPU oracle against a ``qdrant_client`` -- the third-party package
that owns the arithmetic of the reference's retrieval path (reference src/audio_rag/retrieval/qdrant.py:281-332 only
builds its requests).  TEST INFRASTRUCTURE ONLY.

The package is not installable in the build container (no wheel, no network), which is why the oracle is "parity
unpinned".  This module is the pin, ready to run the first time a real ``qdrant-client`` is importable:

    python tests/t0_checklist.py            # real package required; exit code 77 when it is not importable
    python tests/t0_checklist.py --double   # the oracle-backed test double (tests/fake_qdrant): checks THIS file only

It drives ``QdrantClient(":memory:")`` with exactly the calls the reference makes (create_collection :95-118, upsert
:197-220, query_points :281-332) and compares ids bit-exactly / scores within the stated tolerances with
``oracle.OracleIndex``, one function per Appendix B item.  Items whose qdrant behaviour the oracle deliberately does
not copy (root-filter placement in local mode, order of exact ties) are OBSERVED and reported, not asserted.
Run against the double the comparison is circular by construction (the double calls the oracle): it proves the
checklist code works, nothing about qdrant.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FAKE_DIR = os.path.join(HERE, "fake_qdrant")
DIM = 256            # the oracle's canonical dense order needs dim % 256 == 0
DENSE_RTOL = 1e-2    # fp32 cosine of the unrounded vectors vs the canonical bf16 score (north star: 1e-2 rel in bf16)
SPARSE_RTOL = 1e-5   # fp32 of a short fp64 sum vs qdrant's float32 dot product


def _basis(*pairs):
    v = np.zeros(DIM, dtype=np.float32)
    for k, x in pairs:
        v[k] = x
    return v


# Six documents with strictly separated scores in both legs for the query (e0 ; {10: 1, 20: 1}):
#   dense order  : 0 (1.0) 1 (.894) 2 (.707) 3 (.447) 4 (0.0) 5 (-1.0)
#   sparse order : 4 (5.0) 2 (3.0) 3 (2.0) 5 (0.5) 1 (0.0, overlaps on a ZERO weight) ; 0 has no common index
KAT_DENSE = [_basis((0, 1.0)), _basis((0, 1.0), (1, 0.5)), _basis((0, 3.0), (1, 3.0)), _basis((0, 1.0), (1, 2.0)),
             _basis((1, 1.0)), _basis((0, -2.0))]
KAT_SPARSE = [([30], [1.0]), ([10], [0.0]), ([10], [3.0]), ([20, 10], [1.0, 1.0]), ([20], [5.0]), ([10, 40], [0.5, 9.0])]
KAT_Q_DENSE = _basis((0, 0.25))                      # un-normalised on purpose (item 4)
KAT_Q_SPARSE = ([20, 10], [1.0, 1.0])                # unsorted on purpose (R3)


class Harness:
    """One in-memory client + the oracle index fed with the same points (ids = insertion order, R1)."""

    def __init__(self, qc, models):
        self.qc, self.m = qc, models
        self.client = qc.QdrantClient(location=":memory:")
        self.oracles: dict = {}

    def create(self, name, hybrid=True):
        m = self.m
        if hybrid:      # qdrant.py:95-108
            self.client.create_collection(
                collection_name=name,
                vectors_config={"dense": m.VectorParams(size=DIM, distance=m.Distance.COSINE)},
                sparse_vectors_config={"sparse": m.SparseVectorParams(index=m.SparseIndexParams(on_disk=False))})
        else:           # qdrant.py:110-118 (legacy: one unnamed vector)
            self.client.create_collection(collection_name=name,
                                          vectors_config=m.VectorParams(size=DIM, distance=m.Distance.COSINE))
        from oracle import oracle
        self.oracles[name] = {"index": oracle.OracleIndex(DIM), "hybrid": hybrid, "meta": []}

    def add(self, name, dense, sparse=None, metas=None):
        m, o = self.m, self.oracles[name]
        base = o["index"].n
        pts = []
        for i, d in enumerate(dense):
            vec = [float(x) for x in d]
            if o["hybrid"]:
                vec = {"dense": vec}
                if sparse is not None and sparse[i] is not None:
                    vec["sparse"] = m.SparseVector(indices=[int(t) for t in sparse[i][0]],
                                                   values=[float(w) for w in sparse[i][1]])
            pts.append(m.PointStruct(id=base + i, vector=vec,
                                     payload={"text": f"row {base + i}", "metadata": (metas[i] if metas else {})}))
        self.client.upsert(collection_name=name, points=pts)
        o["index"].add_f32(np.asarray(dense, np.float32), sparse if o["hybrid"] else None)
        o["meta"] += list(metas) if metas else [{}] * len(dense)

    # ---- the reference's three requests
    def hybrid(self, name, qd, qs, top_k, flt=None):
        m = self.m
        r = self.client.query_points(
            collection_name=name,
            prefetch=[m.Prefetch(query=[float(x) for x in qd], using="dense", limit=top_k * 2),
                      m.Prefetch(query=m.SparseVector(indices=list(qs[0]), values=list(qs[1])), using="sparse",
                                 limit=top_k * 2)],
            query=m.FusionQuery(fusion=m.Fusion.RRF), limit=top_k, query_filter=self._filter(flt), with_payload=True)
        return [int(p.id) for p in r.points], [p.score for p in r.points]

    def sparse(self, name, qs, top_k, flt=None):
        m = self.m
        r = self.client.query_points(collection_name=name,
                                     query=m.SparseVector(indices=list(qs[0]), values=list(qs[1])), using="sparse",
                                     limit=top_k, query_filter=self._filter(flt), with_payload=True)
        return [int(p.id) for p in r.points], [p.score for p in r.points]

    def dense(self, name, qd, top_k, flt=None, score_threshold=None):
        kw = dict(collection_name=name, query=[float(x) for x in qd], limit=top_k, query_filter=self._filter(flt),
                  with_payload=True)
        if self.oracles[name]["hybrid"]:
            kw["using"] = "dense"
        else:
            kw["score_threshold"] = score_threshold
        r = self.client.query_points(**kw)
        return [int(p.id) for p in r.points], [p.score for p in r.points]

    def _filter(self, flt):
        if not flt:
            return None
        m = self.m
        return m.Filter(must=[m.FieldCondition(key=f"metadata.{k}", match=m.MatchValue(value=v)) for k, v in flt.items()])

    # ---- the oracle's answer to the same request
    def expect(self, name, mode, qd, qs, top_k, flt=None, score_threshold=None):
        from oracle import oracle
        o = self.oracles[name]
        elig = np.ones(o["index"].n, dtype=bool)
        for k, v in (flt or {}).items():
            for r, meta in enumerate(o["meta"]):
                got = meta.get(k)
                elig[r] &= (k in meta) and ((v in got) if isinstance(got, (list, tuple)) else got == v)
        qb = oracle.normalize_bf16(np.asarray(qd, np.float32)[None])[0] if qd is not None else None
        qi, qv = (qs if qs is not None else ([], []))
        ids, sc = o["index"].search(mode, qb, qi, qv, elig, top_k, score_threshold)
        return [int(i) for i in ids], [float(s) for s in sc]


def _kat(h: Harness, name="kat"):
    h.create(name, hybrid=True)
    h.add(name, KAT_DENSE, KAT_SPARSE, metas=[{"lang": "en" if i % 2 == 0 else "de"} for i in range(6)])
    return name


def _close(a, b, rtol):
    return len(a) == len(b) and all(abs(x - y) <= rtol * max(abs(x), abs(y), 1e-30) + 1e-7 for x, y in zip(a, b))


# ------------------------------------------------------------------------------------------ Appendix B, item by item

def item1_rrf_constant(h):
    """RRF = sum of 1/(2 + 0-based position); pos 0 in both legs -> exactly 1.0; only pos 3 of one leg -> 0.2 (R9)."""
    h.create("rrf", hybrid=True)
    # doc 0 is #1 in both legs; doc 4 appears only in the dense leg, at position 3
    dense = [_basis((0, 1.0)), _basis((0, 1.0), (1, 0.5)), _basis((0, 1.0), (1, 1.0)), _basis((0, 1.0), (1, 2.0)),
             _basis((0, 1.0), (1, 1.5))]
    sparse = [([10], [9.0]), ([10], [3.0]), ([10], [2.0]), ([10], [1.0]), ([30], [1.0])]
    h.add("rrf", dense, sparse)
    ids, sc = h.hybrid("rrf", _basis((0, 1.0)), ([10], [1.0]), 5)
    exp_ids, exp_sc = h.expect("rrf", "hybrid", _basis((0, 1.0)), ([10], [1.0]), 5)
    assert ids == exp_ids, (ids, exp_ids)
    assert sc == exp_sc, "fused scores must be bit-equal fp64 sums (R9)"
    assert sc[ids.index(0)] == 1.0 and sc[ids.index(4)] == 0.2
    return {"rrf_k": 2, "ids": ids}


def item2_rrf_tie_order(h):
    """A dense-only candidate at position i ties with the sparse-only one at position i: the dense one first (R10)."""
    h.create("tie", hybrid=True)
    dense = [_basis((0, 1.0)), _basis((0, 1.0), (1, 1.0)), _basis((0, -1.0)), _basis((0, -1.0), (1, -1.0))]
    sparse = [None, None, ([10], [2.0]), ([10], [1.0])]
    h.add("tie", dense, sparse)
    ids, sc = h.hybrid("tie", _basis((0, 1.0)), ([10], [1.0]), 1)      # legs of depth 2: dense [0, 1], sparse [2, 3]
    assert ids == [0] and sc == [0.5], (ids, sc)
    ids, sc = h.hybrid("tie", _basis((0, 1.0)), ([10], [1.0]), 2)      # depth 4: rows 2, 3 sit in both legs now
    assert (ids, sc) == h.expect("tie", "hybrid", _basis((0, 1.0)), ([10], [1.0]), 2)
    return {"order": ids}


def item3_sparse_touched(h):
    """Docs without a common index are absent; a doc overlapping only on a zero weight is present with 0.0 (R7)."""
    name = _kat(h)
    ids, sc = h.sparse(name, KAT_Q_SPARSE, 10)
    assert ids == [4, 2, 3, 5, 1], ids
    assert _close(sc, [5.0, 3.0, 2.0, 0.5, 0.0], SPARSE_RTOL) and sc[-1] == 0.0
    exp = h.expect(name, "sparse", None, KAT_Q_SPARSE, 10)
    assert ids == exp[0] and _close(sc, exp[1], SPARSE_RTOL)
    return {"ids": ids}


def item4_cosine(h):
    """Stored and query vectors are normalised: un-normalised inputs rank like normalised ones (R2)."""
    name = "kat" if "kat" in h.oracles else _kat(h)
    ids, sc = h.dense(name, KAT_Q_DENSE, 6)
    assert ids == [0, 1, 2, 3, 4, 5], ids
    want = [1.0, 2 / 5 ** 0.5, 1 / 2 ** 0.5, 1 / 5 ** 0.5, 0.0, -1.0]
    assert _close(sc, want, DENSE_RTOL), sc
    exp = h.expect(name, "dense", KAT_Q_DENSE, None, 6)
    assert ids == exp[0] and _close(sc, exp[1], DENSE_RTOL)
    return {"scores": sc}


def item5_prefetch_depth(h):
    """Legs have limit 2*top_k and the fused list is cut to top_k: with top_k = 1 a doc that is #2 in both legs wins."""
    h.create("depth", hybrid=True)
    dense = [_basis((0, 1.0)), _basis((0, 1.0), (1, 0.2)), _basis((1, 1.0))]
    sparse = [([30], [1.0]), ([10], [1.0]), ([10], [5.0])]
    h.add("depth", dense, sparse)        # dense leg [0, 1], sparse leg [2, 1]
    ids, sc = h.hybrid("depth", _basis((0, 1.0)), ([10], [1.0]), 1)
    assert ids == [1] and sc == [1.0 / 3 + 1.0 / 3], (ids, sc)
    return {"ids": ids}


def item6_root_filter_on_fusion(h):
    """OBSERVED: where does the root query_filter of a fusion query act in this client (legs / fused list / ignored)?
    The engine applies it inside both legs (R4, server planner semantics)."""
    name = "kat" if "kat" in h.oracles else _kat(h)
    ids, _ = h.hybrid(name, KAT_Q_DENSE, KAT_Q_SPARSE, 1, flt={"lang": "de"})    # 'de' rows: 1, 3, 5
    in_legs = h.expect(name, "hybrid", KAT_Q_DENSE, KAT_Q_SPARSE, 1, flt={"lang": "de"})[0]
    unfiltered = h.expect(name, "hybrid", KAT_Q_DENSE, KAT_Q_SPARSE, 1)[0]
    if ids == in_legs and ids != unfiltered:
        sem = "legs"
    elif ids == unfiltered:
        sem = "ignored"
    else:
        sem = "fused-list"        # filtered after fusion: legs of depth 2 hold rows 0, 1, 4, 2 -> only row 1 survives
    return {"semantics": sem, "ids": ids, "engine_rule": "legs"}


def item7_score_threshold(h):
    """score_threshold on a legacy (unnamed-vector) collection drops lower-scoring hits (R6, qdrant.py:324-332)."""
    h.create("legacy", hybrid=False)
    h.add("legacy", KAT_DENSE)
    ids, sc = h.dense("legacy", KAT_Q_DENSE, 6, score_threshold=0.6)
    assert ids == [0, 1, 2], (ids, sc)
    assert ids == h.expect("legacy", "dense", KAT_Q_DENSE, None, 6, score_threshold=0.6)[0]
    return {"ids": ids}


def item8_score_types(h):
    """OBSERVED: python types of hit.score per request kind."""
    name = "kat" if "kat" in h.oracles else _kat(h)
    return {"dense": type(h.dense(name, KAT_Q_DENSE, 1)[1][0]).__name__,
            "sparse": type(h.sparse(name, KAT_Q_SPARSE, 1)[1][0]).__name__,
            "rrf": type(h.hybrid(name, KAT_Q_DENSE, KAT_Q_SPARSE, 1)[1][0]).__name__}


def item9_duplicate_tie_order(h):
    """OBSERVED: order of exactly equal dense scores (duplicate vectors).  The engine's rule R5 is smaller row id
    first; qdrant-local's is whatever its argsort yields, so such cases are excluded from id comparisons."""
    h.create("dup", hybrid=False)
    v = _basis((0, 1.0), (3, 2.0))
    h.add("dup", [v, _basis((5, 1.0)), v, v * 3.0])
    ids, sc = h.dense("dup", v, 3)
    assert sorted(ids) == [0, 2, 3] and max(sc) - min(sc) <= 1e-6
    return {"order": ids, "engine_rule": [0, 2, 3]}


def item10_empty_collection(h):
    """Searching an empty, just-created collection returns [] (the plugin creates unknown names, qdrant.py:248)."""
    h.create("empty", hybrid=True)
    assert h.hybrid("empty", KAT_Q_DENSE, KAT_Q_SPARSE, 3)[0] == []
    assert h.dense("empty", KAT_Q_DENSE, 3)[0] == []
    return {}


def random_differential(h, n=300, nq=8, seed=5):
    """Seeded planted-query corpus: ids of all three request kinds equal the oracle's; scores within tolerance."""
    sys.path.insert(0, os.path.join(ROOT, "audio-rag_b200"))
    from b200rag import synth
    V = 250_002
    thr = synth.zipf_thresholds(V)
    f = synth.dense_rows_f32(seed, 0, n, DIM)
    ip, tt, ww = synth.sparse_docs_csr(seed, 0, n, 10_000, V, 64, thr, synth.bm25_tables(10_000, V, 64))
    sparse = [([int(t) for t in tt[ip[i]:ip[i + 1]]], [float(w) for w in ww[ip[i]:ip[i + 1]]]) for i in range(n)]
    h.create("rand", hybrid=True)
    h.add("rand", f, sparse, metas=[{"lang": "en" if i % 3 else "de"} for i in range(n)])
    qf = synth.dense_queries_f32(seed + 1, 0, nq, n, DIM, corpus_seed=seed)
    qi, qt, qw = synth.sparse_queries(seed + 1, 0, nq, 10, V, thr)
    checked = skipped = 0
    for i in range(nq):
        if i % 10 == 9:          # pure-random queries: near-uniform scores, bf16 rounding may reorder them
            continue
        qs = ([int(t) for t in qt[qi[i]:qi[i + 1]]], [float(w) for w in qw[qi[i]:qi[i + 1]]])
        for k in (3, 10):
            ids, sc = h.dense("rand", qf[i], k)
            ei, es = h.expect("rand", "dense", qf[i], None, k)
            gaps = -np.diff(np.asarray(h.expect("rand", "dense", qf[i], None, k + 1)[1]))
            if len(gaps) and gaps.min() < 2e-3:     # closer than the bf16 storage error: order not comparable
                skipped += 1
            else:
                assert ids == ei, ("dense", i, k, ids, ei)
                assert _close(sc, es, DENSE_RTOL)
                checked += 1
            ids, sc = h.sparse("rand", qs, k)
            ei, es = h.expect("rand", "sparse", None, qs, k)
            if len(set(es)) == len(es):
                assert ids == ei, ("sparse", i, k, ids, ei)
                assert _close(sc, es, SPARSE_RTOL)
                checked += 1
            else:
                skipped += 1
    return {"checked": checked, "skipped_near_ties": skipped}


ITEMS = [item1_rrf_constant, item2_rrf_tie_order, item3_sparse_touched, item4_cosine, item5_prefetch_depth,
         item6_root_filter_on_fusion, item7_score_threshold, item8_score_types, item9_duplicate_tie_order,
         item10_empty_collection, random_differential]


def load_client(double: bool):
    """(qdrant_client module, models module, version string) or None.  ``double`` selects tests/fake_qdrant."""
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    if double:
        if FAKE_DIR not in sys.path:
            sys.path.insert(0, FAKE_DIR)
    else:
        sys.path[:] = [p for p in sys.path if os.path.abspath(p) != FAKE_DIR]
        for extra in (os.path.join(ROOT, "baseline", "_ref"),):     # where a later round may drop the real package
            if os.path.isdir(extra) and extra not in sys.path:
                sys.path.append(extra)
    try:
        import qdrant_client
        from qdrant_client import models
    except ImportError:
        return None
    is_double = os.path.abspath(getattr(qdrant_client, "__file__", "")).startswith(FAKE_DIR)
    if is_double != double:
        return None
    version = "test double"
    if not double:
        try:
            from importlib.metadata import version as _v
            version = _v("qdrant-client")
        except Exception:
            version = "unknown"
    return qdrant_client, models, version


def probe() -> dict:
    """Where a real ``qdrant_client`` was looked for and what was found (run on the GPU pod by
    tests/test_oracle_vs_qdrant.py::test_t0_probe_on_gpu_pod, so the outcome is on record where a wheel could exist)."""
    import glob
    import importlib.util
    sys.path[:] = [p for p in sys.path if os.path.abspath(p) != FAKE_DIR]
    searched = [os.path.join(ROOT, "baseline", "_ref")] + [p for p in sys.path if p]
    for extra in (os.path.join(ROOT, "baseline", "_ref"),):
        if os.path.isdir(extra) and extra not in sys.path:
            sys.path.append(extra)
    wheels = []
    for d in ("/opt/wheelhouse", os.path.join(ROOT, "wheelhouse"), os.path.join(ROOT, "baseline"), "/root/wheelhouse", "/tmp"):
        if os.path.isdir(d):
            wheels += glob.glob(os.path.join(d, "qdrant*"))[:20] + glob.glob(os.path.join(d, "*", "qdrant*"))[:20]
    out = {"python": sys.version.split()[0], "searched_sys_path": searched, "searched_wheel_dirs":
           ["/opt/wheelhouse", "<repo>/wheelhouse", "<repo>/baseline", "/root/wheelhouse", "/tmp"], "wheels_found": wheels}
    try:
        spec = importlib.util.find_spec("qdrant_client")
    except Exception as e:                      # a broken install can raise here
        spec, out["find_spec_error"] = None, repr(e)
    if spec is None:
        out.update({"found": False, "error": "ModuleNotFoundError: No module named 'qdrant_client'"})
        return out
    try:
        import qdrant_client  # noqa: F401
        from importlib.metadata import version as _v
        out.update({"found": True, "version": _v("qdrant-client"), "location": os.path.dirname(qdrant_client.__file__)})
    except Exception as e:
        out.update({"found": False, "error": repr(e)})
    return out


def run(double: bool) -> dict | None:
    loaded = load_client(double)
    if loaded is None:
        return None
    qc, models, version = loaded
    report = {"client": version, "pins_oracle": not double, "items": {}}
    h = Harness(qc, models)
    for fn in ITEMS:
        report["items"][fn.__name__] = fn(h)
    return report


if __name__ == "__main__":
    if "--probe" in sys.argv:
        print(json.dumps(probe()))
        sys.exit(0)
    rep = run(double="--double" in sys.argv)
    if rep is None:
        print("qdrant_client is not importable: oracle stays PARITY UNPINNED", file=sys.stderr)
        sys.exit(77)
    print(json.dumps(rep))
