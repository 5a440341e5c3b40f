"""b200rag: B200-native hybrid retrieval engine behind audio-rag's `src/audio_rag/retrieval` interface.

Host side (Python, like the reference) over a C-ABI CUDA library (include/b200rag.h):
  _ffi.py       ctypes binding + `Shard`, `ShardGroup` (several shards driven by one process)
  retriever.py  `B200Retriever` -- mirror of `QdrantRetriever` (reference src/audio_rag/retrieval/qdrant.py)
  dist.py       one-process-per-GPU row sharding, NCCL all-gather of per-shard candidates
  synth.py      deterministic synthetic corpus/query generators (twins of csrc/synth.cu)
"""
from . import _ffi, synth  # noqa: F401
from ._ffi import B200RagError, Shard, ShardGroup, device_count, normalize_bf16  # noqa: F401

__all__ = ["B200RagError", "Shard", "ShardGroup", "device_count", "normalize_bf16", "synth"]
