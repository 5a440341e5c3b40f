"""A model of the peer-memory candidate exchange under adversarial scheduling -- no GPU, no CUDA.

What is modelled (the ordering rules of b200rag/dist.py::ShardedSearcher.run_staged and csrc/select.cu's
exchange_kernel / fuse_kernel, nothing of their arithmetic):

  * every rank runs searches 0..S-1; search i has epoch i+1 and window parity (i+1) & 1 (engine.cu: ++x_epoch, then
    parity = x_epoch & 1);
  * LEGS(i) writes the rank's candidate block `mine`; EXCH(i) stores it into slot [parity][rank] of EVERY rank's
    window and then publishes epoch i+1 in that rank's flag word; FUSE(i) starts once all `world` flags of the local
    window carry an epoch >= i+1, reads all slots of that parity over an interval, and ends;
  * serial mode: LEGS, EXCH, FUSE of a rank run in stream order on ONE stream;
  * pipelined mode (B200RAG_PIPELINE_TAIL=1): LEGS on the main stream, EXCH + FUSE on the tail stream; EXCH(i) waits
    for LEGS(i) (legs_done event); `mine` and the result block are double-buffered by search parity and LEGS(i) waits
    for FUSE(i-2) (tail_done event).

A random scheduler picks any enabled operation of any rank, thousands of times per seed.  Checked: no deadlock (some
operation is always enabled until every rank finished), every FUSE reads exactly the blocks of ITS search from every
rank (no stale slot, no slot overwritten while being read), and LEGS never overwrites a block its own EXCH has not
shipped yet.  Mutants of the protocol (one parity, no tail_done wait, flags published before the data) must be caught
by the same checks -- otherwise the model would prove nothing."""
import random

import pytest


class Rank:
    def __init__(self, world, pipelined):
        self.main_pc = 0          # next LEGS index
        self.tail_pc = 0          # next tail op: 2*i = EXCH(i), 2*i+1 = FUSE(i)
        self.fuse_open = None     # (search, snapshot) while a FUSE is reading
        self.fused = -1           # last completed FUSE
        self.mine = {}            # buffer parity -> (search) token written by LEGS
        self.mine_shipped = {}    # buffer parity -> search whose EXCH has read it
        self.window = {}          # (parity, src_rank) -> token (src, search)
        self.flags = [0] * world  # epoch published by each source rank
        self.pipelined = pipelined


class Model:
    def __init__(self, world, searches, pipelined, parities=2, wait_tail_done=True, flag_first=False, seed=0):
        self.world, self.S, self.pipelined = world, searches, pipelined
        self.parities, self.wait_tail_done, self.flag_first = parities, wait_tail_done, flag_first
        self.rng = random.Random(seed)
        self.ranks = [Rank(world, pipelined) for _ in range(world)]
        self.errors = []
        self.half_done = {}       # (rank, search) -> {peer: stores done} (EXCH is neither atomic across peers nor per peer)

    # ---- enabledness
    def _buf(self, i):
        return i & 1 if self.pipelined else 0

    def _legs_enabled(self, r):
        k = self.ranks[r]
        i = k.main_pc
        if i >= self.S:
            return False
        if not self.pipelined:
            return k.tail_pc == 2 * i                      # one stream: after FUSE(i-1)
        if self.wait_tail_done and i >= 2 and k.fused < i - 2:
            return False
        return True

    def _tail_enabled(self, r):
        k = self.ranks[r]
        if k.tail_pc >= 2 * self.S:
            return False
        i, is_fuse = divmod(k.tail_pc, 2)
        if not is_fuse:
            return k.main_pc > i                           # legs_done(i)
        if k.fuse_open is not None:
            return True                                    # finishing the read is always possible
        return all(f >= i + 1 for f in k.flags)            # the kernel spins until then

    def enabled(self):
        ops = []
        for r in range(self.world):
            if self._legs_enabled(r):
                ops.append((r, "main"))
            if self._tail_enabled(r):
                ops.append((r, "tail"))
        return ops

    # ---- transitions
    def step(self, r, which):
        k = self.ranks[r]
        if which == "main":
            i = k.main_pc
            b = self._buf(i)
            if b in k.mine and k.mine_shipped.get(b, -1) < k.mine[b]:
                self.errors.append(f"rank {r}: LEGS({i}) overwrites block of search {k.mine[b]} before its EXCH read it")
            k.mine[b] = i
            k.main_pc += 1
            return
        i, is_fuse = divmod(k.tail_pc, 2)
        parity = (i + 1) % self.parities
        if not is_fuse:
            # EXCH(i): one store per scheduler step (the kernel's blocks are independent): per peer the data, then the flag
            stage = self.half_done.setdefault((r, i), {p: 0 for p in range(self.world)})
            p = self.rng.choice([p for p, st in stage.items() if st < 2])
            dst = self.ranks[p]
            write_data = (stage[p] == 0) != self.flag_first          # the mutant publishes the flag first
            if write_data:
                if dst.fuse_open is not None and (dst.fuse_open[0] + 1) % self.parities == parity:
                    self.errors.append(f"rank {r}: EXCH({i}) writes rank {p}'s parity-{parity} slot while "
                                       f"FUSE({dst.fuse_open[0]}) reads it")
                dst.window[(parity, r)] = (r, k.mine[self._buf(i)])
            else:
                dst.flags[r] = i + 1
            stage[p] += 1
            if all(st == 2 for st in stage.values()):
                k.mine_shipped[self._buf(i)] = i
                k.tail_pc += 1
            return
        if k.fuse_open is None:                            # FUSE(i) begins: snapshot what it reads
            snap = [k.window.get((parity, s)) for s in range(self.world)]
            k.fuse_open = (i, snap)
            return
        i0, snap = k.fuse_open                             # FUSE(i) ends: the slots must not have changed, and be search i's
        now = [k.window.get((parity, s)) for s in range(self.world)]
        want = [(s, i0) for s in range(self.world)]
        if snap != want or now != want:
            self.errors.append(f"rank {r}: FUSE({i0}) read {snap} .. {now}, wanted {want}")
        k.fuse_open = None
        k.fused = i0
        k.tail_pc += 1

    def run(self, max_steps=200_000):
        for _ in range(max_steps):
            ops = self.enabled()
            if not ops:
                break
            self.step(*self.rng.choice(ops))
        finished = all(k.main_pc == self.S and k.tail_pc == 2 * self.S for k in self.ranks)
        if not finished:
            self.errors.append("deadlock: no operation enabled before every rank finished")
        return self.errors


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("pipelined", [False, True])
def test_protocol_is_safe_and_live(world, pipelined):
    for seed in range(40):
        errs = Model(world, searches=14, pipelined=pipelined, seed=seed).run()
        assert errs == [], (world, pipelined, seed, errs[:3])


def _caught(**mutant):
    return any(Model(w, searches=14, seed=seed, **mutant).run() for w in (2, 4) for seed in range(40))


def test_mutants_are_caught():
    # one slot parity: a rank one search ahead overwrites what a slower peer's fuse is still reading / has not read yet
    assert _caught(pipelined=False, parities=1)
    assert _caught(pipelined=True, parities=1)
    # pipelined without the tail_done wait: LEGS(i) may overwrite the block EXCH(i-2) has not shipped
    assert _caught(pipelined=True, wait_tail_done=False)
    # epoch flag published before the data: a fuse may start on a half-written slot
    assert _caught(pipelined=False, flag_first=True)
