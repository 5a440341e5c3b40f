"""Minimal stand-ins for the qdrant_client.models names the reference imports (qdrant.py:81-84, 167, 253-257)."""
from dataclasses import dataclass, field
from enum import Enum
from typing import Any


class Distance(str, Enum):
    COSINE = "Cosine"
    DOT = "Dot"
    EUCLID = "Euclid"


class Fusion(str, Enum):
    RRF = "rrf"
    DBSF = "dbsf"


@dataclass
class VectorParams:
    size: int
    distance: Distance


@dataclass
class SparseIndexParams:
    on_disk: bool = False


@dataclass
class SparseVectorParams:
    index: SparseIndexParams | None = None
    modifier: Any = None


@dataclass
class SparseVector:
    indices: list
    values: list


@dataclass
class PointStruct:
    id: Any
    vector: Any
    payload: dict | None = None


@dataclass
class MatchValue:
    value: Any


@dataclass
class FieldCondition:
    key: str
    match: MatchValue


@dataclass
class Filter:
    must: list = field(default_factory=list)


@dataclass
class Prefetch:
    query: Any
    using: str | None = None
    limit: int = 10
    filter: Filter | None = None


@dataclass
class FusionQuery:
    fusion: Fusion


@dataclass
class ScoredPoint:
    id: Any
    score: float
    payload: dict | None = None
    version: int = 0
