"""Generates tests/golden/retrieval_golden.json: expected outputs of the REFERENCE's own plugin code
(`QdrantRetriever`, /root/reference/src/audio_rag/retrieval/qdrant.py) driven against the oracle-backed qdrant
test double, for fixed seeded inputs (tests/data_small.py).  The fixture travels to the GPU box, where
/root/reference does not exist, and pins B200Retriever's results there (tests/test_gpu_retriever.py).

    python tests/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import conftest  # noqa: E402,F401  (wires sys.path: reference, test doubles, package)
from data_small import DIM, make_chunks, make_queries  # noqa: E402

CASES = [
    # (search_type, top_k, collection, filter)
    ("hybrid", 5, "tenant_a", None), ("hybrid", 10, "tenant_b", None), ("dense", 5, "tenant_a", None),
    ("sparse", 5, "tenant_a", None), ("hybrid", 5, "legacy", None), ("sparse", 3, "legacy", None),
    ("hybrid", 5, "tenant_a", {"lang": "de"}), ("dense", 8, "tenant_b", {"tags": "a"}),
    ("hybrid", 50, "tenant_b", None), ("hybrid", 5, "nope", None),
]


def corpus(types):
    A, E, S = types
    return {"tenant_a": make_chunks(120, 11, "A", A, E, S), "tenant_b": make_chunks(80, 12, "B", A, E, S),
            "legacy": make_chunks(60, 13, "L", A, E, S, sparse=False)}


def load(retriever, types):
    data = corpus(types)
    for part in range(2):
        for name, (ch, em) in data.items():
            h = len(ch) // 2
            sl = slice(0, h) if part == 0 else slice(h, None)
            retriever.add(ch[sl], em[sl], name)
    return data


def run_cases(retriever, types):
    A, E, S = types
    qs = make_queries(3, 21, 120, 11, E, S)
    out = []
    for st, k, name, flt in CASES:
        for qi, q in enumerate(qs):
            res = retriever.search(q, top_k=k, collection_name=name, filter_metadata=flt, search_type=st)
            out.append({"search_type": st, "top_k": k, "collection": name, "filter": flt, "query": qi,
                        "texts": [r.chunk.text for r in res], "scores": [float(r.score).hex() for r in res]})
    return out


def generate():
    from audio_rag.config import RetrievalConfig
    from audio_rag.core import AudioChunk, EmbeddingResult, SparseVector
    from audio_rag.retrieval import RetrievalRegistry
    types = (AudioChunk, EmbeddingResult, SparseVector)
    ref = RetrievalRegistry.create("qdrant", config=RetrievalConfig(qdrant_in_memory=True), embedding_dim=DIM)
    load(ref, types)
    return {"dim": DIM, "generator": "tests/make_golden.py", "source": "reference QdrantRetriever + tests/fake_qdrant",
            "cases": run_cases(ref, types)}


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "golden"), exist_ok=True)
    path = os.path.join(HERE, "golden", "retrieval_golden.json")
    json.dump(generate(), open(path, "w"), indent=0)
    print(path, os.path.getsize(path))
