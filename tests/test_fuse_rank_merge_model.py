"""Host model of fuse_kernel's merge of sorted shard blocks (audio-rag_b200/csrc/select.cu): every shard's L candidates of a
leg arrive in leg order (valid first, score descending, id ascending); a candidate's merged rank is its own position
plus, for every OTHER shard, the number of that shard's valid candidates that are better than it -- found by a binary
search with exactly the kernel's probe (`o.valid && cand_better(o, e)` -> go right).  The model checks that these
ranks are a permutation of 0..n_valid-1 equal to the global sort under the same order, for random blocks with score ties
across shards, short blocks (fewer than L valid) and empty shards; and that `block_unsorted` flags exactly the blocks a
foreign caller shuffled.  CPU only; the kernel is compared with a single shard and the oracle in the GPU tests."""
import numpy as np
from hypothesis import given, settings, strategies as st


def better(sa, ia, sb, ib):
    """cand_better: score (as ordered fp32) descending, then id ascending."""
    return sa > sb or (sa == sb and ia < ib)


def merged_ranks(blocks, L):
    """blocks: per shard a list of (valid, score, id) of length L, in leg order."""
    ranks = {}
    for sh, blk in enumerate(blocks):
        for j, (v, s, i) in enumerate(blk):
            if not v:
                continue
            rank = j
            for b, other in enumerate(blocks):
                if b == sh:
                    continue
                lo, hi = 0, L
                while lo < hi:
                    mid = (lo + hi) >> 1
                    ov, os_, oi = other[mid]
                    if ov and better(os_, oi, s, i):
                        lo = mid + 1
                    else:
                        hi = mid
                rank += lo
            ranks[(sh, j)] = rank
    return ranks


def block_unsorted(blk):
    for j in range(1, len(blk)):
        v, s, i = blk[j]
        pv, ps, pi = blk[j - 1]
        if v and not (pv and better(ps, pi, s, i)):
            return True
    return False


def make_blocks(rng, n_shards, L, tie_levels):
    ids = rng.permutation(n_shards * L * 4)[: n_shards * L]            # globally unique ids
    blocks, k = [], 0
    for sh in range(n_shards):
        nvalid = int(rng.integers(0, L + 1))
        scores = rng.integers(0, tie_levels, nvalid).astype(np.float32) / np.float32(tie_levels)   # many exact ties
        entries = sorted(((float(s), int(ids[k + t])) for t, s in enumerate(scores)), key=lambda e: (-e[0], e[1]))
        k += L
        blocks.append([(True, s, i) for s, i in entries] + [(False, 0.0, -1)] * (L - nvalid))
    return blocks


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 2 ** 32 - 1), st.integers(1, 8), st.integers(1, 40), st.sampled_from([2, 5, 1000]))
def test_rank_merge_equals_global_sort(seed, n_shards, L, tie_levels):
    rng = np.random.default_rng(seed)
    blocks = make_blocks(rng, n_shards, L, tie_levels)
    assert not any(block_unsorted(b) for b in blocks)
    ranks = merged_ranks(blocks, L)
    flat = [(s, i) for blk in blocks for (v, s, i) in blk if v]
    order = sorted(flat, key=lambda e: (-e[0], e[1]))
    assert sorted(ranks.values()) == list(range(len(flat)))           # a permutation: no slot written twice, none left out
    for (sh, j), r in ranks.items():
        _, s, i = blocks[sh][j]
        assert order[r] == (s, i)


def test_unsorted_blocks_are_detected():
    rng = np.random.default_rng(1)
    for _ in range(200):
        blocks = make_blocks(rng, 3, 12, 5)
        blk = list(blocks[1])
        nvalid = sum(1 for e in blk if e[0])
        if nvalid < 2:
            continue
        a, b = rng.choice(nvalid, 2, replace=False)
        blk[a], blk[b] = blk[b], blk[a]                                # two distinct candidates swapped: order broken
        assert block_unsorted(blk)
        hole = list(blocks[1])
        hole[0] = (False, 0.0, -1)                                     # an invalid entry ahead of valid ones
        assert block_unsorted(hole)
