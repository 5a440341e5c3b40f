"""Row-sharded search: one process per GPU, `torch.distributed` for the plumbing (NCCL over NVLink on GPUs).

The path shards naturally: rows are independent in both legs, so rank r holds the contiguous row
range [base_r, base_r + n_r) as its own `Shard` (global id = row_base + local) and answers every query for its
rows.  The ONE exchange step is an all-gather of the per-shard candidate lists (already exact-scored and ordered
under R5): [nlegs, B, L] 16-byte candidates + one trailer per rank, <= 1.6 MB per rank at B = 1024, L = 200 and
656 bytes at B = 1, L = 20.  Every rank then runs the same merge + RRF kernel on the gathered buffer, so the
fused result is identical on all ranks and bit-identical to a single-shard search (RRF needs GLOBAL ranks, so it
runs after the gather, never per shard).

On one box the exchange does not go through a collective at all: every rank exports an exchange window over CUDA
IPC, `b200rag_p2p_exchange` stores the rank's block into all peers' windows with NVLink peer stores and publishes an
epoch flag, and the fuse kernel itself waits for the flags (include/b200rag.h, "peer-memory candidate exchange").
The NCCL all-gather stays as the fallback (B200RAG_P2P=0, blocks larger than the window slot, non-CUDA test doubles).

torch is plumbing only: it owns the exchange buffers and the process group; all compute is `libb200rag.so`.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from ._ffi import Shard


PIPELINE_TAIL_DEFAULT = "1"


def shard_bounds(n_total: int, world: int, rank: int, align: int = 1) -> tuple[int, int]:
    """Contiguous row range of `rank` (balanced; starts aligned to `align` rows)."""
    per = -(-n_total // world)
    per = -(-per // align) * align
    lo = min(n_total, rank * per)
    hi = min(n_total, lo + per)
    return lo, hi


class ShardedSearcher:
    """One rank's end of a sharded corpus.  `shard` is this rank's `Shard` (or a duck-typed double in CPU tests)."""

    def __init__(self, shard, device: torch.device, group=None):
        self.shard = shard
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._bufs: dict = {}
        self._slots: dict = {}
        self.broken = None          # reason, once an exchange timed out (the ranks' epochs have diverged: reset())
        # Everything this class enqueues besides the library's kernels -- the buffers' zero fills, the NCCL all-gather,
        # the result read-back, the pipelined tail's events -- runs on torch's CURRENT stream.  A Shard runs on a
        # private stream until told otherwise, which would leave the two unordered (results read before the fuse
        # finished, candidates gathered before the legs).  So the shard is tied to torch's current stream here.
        if device.type == "cuda" and hasattr(shard, "set_stream"):
            shard.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.exact_fallback = os.environ.get("B200RAG_EXACT_FALLBACK", "1") != "0"
        self.p2p = False
        self.p2p_slot_bytes = int(os.environ.get("B200RAG_P2P_SLOT_BYTES", 8 << 20))
        if self.world > 1 and device.type == "cuda" and hasattr(shard, "p2p_export") and \
                os.environ.get("B200RAG_P2P", "1") != "0":
            self._setup_p2p()
        # Pipelined searches (b200rag_set_pipeline): only the dense scan of a search stays on the shard's stream; the
        # sparse leg, both legs' tails, the exchange and the fuse run on the library's result stream, so the NEXT
        # search's scan starts the moment this one's ends -- also while this search still waits for its slowest peer.
        # Everything that touches the candidate and result buffers is on that one in-order stream, so single buffers
        # suffice; results are read there (fetch / result_stream).  B200RAG_PIPELINE_TAIL=0/1.
        self.pipeline = bool((self.p2p or self.world == 1) and device.type == "cuda" and hasattr(shard, "set_pipeline")
                             and os.environ.get("B200RAG_PIPELINE_TAIL", PIPELINE_TAIL_DEFAULT) == "1")
        self._tail = None
        self._paused = False
        if self.pipeline:
            # a torch-owned stream handed to the library (torch ops on a foreign stream would outlive it at teardown)
            self._tail = torch.cuda.Stream(device=device)
            self.shard.set_pipeline(True, self._tail.cuda_stream)

    def _setup_p2p(self):
        """Exchange CUDA IPC handles of the per-rank windows; every rank must succeed or all fall back to NCCL."""
        ok, handle = 1, b""
        try:
            handle = self.shard.p2p_export(self.world, self.p2p_slot_bytes)
        except Exception:
            ok = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, handle), group=self.group)
        if all(h[0] for h in handles):
            try:
                self.shard.p2p_attach(self.rank, self.world, b"".join(h[1] for h in handles))
            except Exception:
                ok = 0
        else:
            ok = 0
        oks = [None] * self.world
        dist.all_gather_object(oks, ok, group=self.group)
        self.p2p = all(oks)
        if not self.p2p and ok:
            try:
                self.shard.p2p_close()
            except Exception:
                pass

    def _buffers(self, nlegs, B, L, k):
        key = (nlegs, B, L, k)
        b = self._bufs.get(key)
        if b is None:
            n = nlegs * B * L + 1                              # + trailer (ambiguity counter)
            b = {
                "mine": torch.zeros((n, 2), dtype=torch.int64, device=self.device),
                "all": torch.zeros((self.world, n, 2), dtype=torch.int64, device=self.device),
                # ids | scores(f64 bits) | counts [B] + ambiguity counter + sticky timeout latch, in ONE buffer -> one
                # D2H per batch (zeroed once here: the latch is only ever set by a fuse that gave up on a peer)
                "out": torch.zeros(2 * B * k + (B + 3) // 2 + 1, dtype=torch.int64, device=self.device),
            }
            b["host"] = torch.empty_like(b["out"], device="cpu")
            if self.device.type == "cuda":
                b["host"] = b["host"].pin_memory()
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()      # the zero fills precede any use on any stream
            self._bufs[key] = b
        return b

    def result_stream(self):
        """The stream on which a search's fused results become available (enqueue dependent work there)."""
        if self._tail is not None:
            return self._tail
        return torch.cuda.current_stream(self.device) if self.device.type == "cuda" else None

    def record_end(self, event):
        """Record a (timing) event at the point where the last enqueued search is complete."""
        event.record(self.result_stream())

    def broadcast_query(self, arrays: dict | None, src: int = 0) -> dict:
        """Replicate a query batch from `src` (serving: the rank that received the request) to all ranks."""
        if self.world == 1:
            return arrays
        obj = [arrays if self.rank == src else None]
        dist.broadcast_object_list(obj, src=src, group=self.group)
        return obj[0]

    def stage(self, mode, top_k, q_bits=None, sp_indptr=None, sp_terms=None, sp_weights=None, mask_ids=None,
              score_threshold=None, rrf_k=0, slot=None):
        """Copy a query batch to the device.  With `slot` the batch gets its own device block and stays resident:
        `use_slot(slot)` re-activates it later without a copy (a queue of batches enqueued back to back)."""
        q, keep = self.shard.make_query(mode, top_k, q_bits, sp_indptr, sp_terms, sp_weights, mask_ids,
                                        score_threshold, rrf_k)
        if slot is None:
            self.shard.stage(q, keep)
        else:
            self.shard.stage(q, keep, slot=slot)
        nlegs, L = Shard.legs_len(q)
        self._cur = (nlegs, q.batch, L, top_k)
        if slot is not None:
            self._slots[slot] = self._cur
        return self._cur

    def use_slot(self, slot):
        self.shard.use_slot(slot)
        self._cur = self._slots[slot]
        return self._cur

    def run_staged(self):
        """legs -> all-gather -> fuse on the device; results stay in the device out buffer (no host sync)."""
        nlegs, B, L, k = self._cur
        b = self._buffers(nlegs, B, L, k)
        mine, allb, out = b["mine"], b["all"], b["out"]
        if self.pipeline and not self._paused and (self.world == 1 or mine.numel() * 8 <= self.p2p_slot_bytes):
            self.shard.legs(mine, mine[-1])                # scan on the shard's stream, tails on the result stream
            if self.world > 1:
                self.shard.p2p_exchange(mine, mine.numel() * 8)          # (the library enqueues these on the result
                self.shard.p2p_fuse(out[:B * k], out[B * k:2 * B * k], out[2 * B * k:])     #  stream, after the tails)
            else:
                self.shard.fuse(mine, 1, out[:B * k], out[B * k:2 * B * k], out[2 * B * k:], has_trailer=True)
            return {"out": out, "host": b["host"], "tail": True}
        self.shard.legs(mine, mine[-1])                    # (legs zeroes the trailer's ambiguity counter itself)
        if self.world > 1 and self.p2p and mine.numel() * 8 <= self.p2p_slot_bytes:
            # peer stores into every rank's window + epoch flags; the fuse kernel waits for the flags itself
            self.shard.p2p_exchange(mine, mine.numel() * 8)
            self.shard.p2p_fuse(out[:B * k], out[B * k:2 * B * k], out[2 * B * k:])
            return b
        if self.world > 1:
            # output as the concatenation along dim 0 (the layout both NCCL and gloo accept)
            dist.all_gather_into_tensor(allb.view(-1, 2), mine, group=self.group)
            src = allb
        else:
            src = mine
        self.shard.fuse(src, self.world, out[:B * k], out[B * k:2 * B * k], out[2 * B * k:], has_trailer=True)
        return b

    def fetch(self, b):
        """Device -> pinned host read-back of the fused results: (ids [B,k], scores [B,k] f64, counts [B], ambiguous)."""
        nlegs, B, L, k = self._cur
        if b.get("tail"):
            with torch.cuda.stream(self._tail):
                b["host"].copy_(b["out"], non_blocking=True)
            self._tail.synchronize()
        else:
            b["host"].copy_(b["out"], non_blocking=True)
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()
        h = b["host"].numpy()
        ids = h[:B * k].reshape(B, k).copy()
        scores = h[B * k:2 * B * k].view(np.float64).reshape(B, k).copy()
        cnt = h[2 * B * k:].view(np.int32)
        amb = -1 if cnt[B + 1] != 0 else int(cnt[B])      # the latch: some block of some fuse gave up on a peer
        return ids, scores, cnt[:B].copy(), amb

    def search(self, mode, top_k, q_bits=None, sp_indptr=None, sp_terms=None, sp_weights=None, mask_ids=None,
               score_threshold=None, rrf_k=0, max_retries=4):
        """Whole sharded path with HOST buffers (query replicated on every rank).  Returns (ids, scores, counts).

        Exactness: when some shard's slack guard flags the result, every rank widens its slack and repeats (the
        ambiguity counter is global, so all ranks take the same decision); after `max_retries` the legs are recomputed
        EXHAUSTIVELY (always exact) -- or, with B200RAG_EXACT_FALLBACK=0, the search raises instead of returning a
        result that could differ from the exact top-k."""
        if self.broken:
            raise RuntimeError(f"sharded search: {self.broken}; call reset() on every rank")
        nlegs, B, L, k = self.stage(mode, top_k, q_bits, sp_indptr, sp_terms, sp_weights, mask_ids, score_threshold,
                                    rrf_k)
        slack0 = None
        exhaustive = False
        if self.pipeline:                 # a lone synchronous search: classic form, everything on the shard's stream
            self._paused = True
            self.shard.pipeline_pause(True)
        try:
            for attempt in range(max_retries + 2):
                ids, scores, counts, amb = self.fetch(self.run_staged())
                if amb < 0:
                    self.broken = "a peer's candidates never arrived (exchange flag timed out)"
                    raise RuntimeError(f"sharded search: {self.broken}")
                if amb == 0:
                    break
                if exhaustive:
                    raise RuntimeError("sharded search: the exhaustive pass reported ambiguity (internal error)")
                if attempt >= max_retries:
                    if not self.exact_fallback:
                        raise RuntimeError("sharded search: the slack guard never cleared (ties or near-duplicate "
                                           "scores around the top-k cut) and the exhaustive exact pass is disabled")
                    exhaustive = True
                    self.shard.set_exhaustive(True)
                    continue
                # some shard's slack guard failed: widen on every rank (same decision everywhere: amb is global)
                slack0 = max(16, L // 2) if slack0 is None else slack0
                slack0 = slack0 * 2 + L
                self.shard.set_slack(min(slack0, 3 * 256 - L))
        finally:
            if slack0 is not None:
                self.shard.set_slack(0)
            if exhaustive:
                self.shard.set_exhaustive(False)
            if self._paused:
                self._paused = False
                self.shard.pipeline_pause(False)
        return ids, scores, counts

    def reset(self):
        """Collective: tear the peer windows down and set them up again (after an exchange timeout the ranks' epochs
        and parities no longer agree).  Every rank must call it."""
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        if self.p2p:
            try:
                self.shard.p2p_close()
            except Exception:
                pass
        self._bufs.clear()
        self.p2p = False
        if self.world > 1 and self.device.type == "cuda" and hasattr(self.shard, "p2p_export") and \
                os.environ.get("B200RAG_P2P", "1") != "0":
            self._setup_p2p()
        if self.pipeline and not (self.p2p or self.world == 1):
            self.pipeline = False
            self.shard.set_pipeline(False)
            self._tail = None
        self.broken = None
