"""A model of the in-library pipelined searches (csrc/engine.cu::run_legs, piped form) under adversarial
scheduling -- no GPU, no CUDA.  Only the ORDERING rules are modelled, nothing of the arithmetic:

  main stream, search i (parity p = i & 1):   wait ev_tail[p] (recorded by search i-2)  ->  memset thr[p]  ->  record ev_fork
                                              ->  SCAN(i): writes lists[p], raises thr[p]  ->  record ev_scan[p]
  second stream, search i:                    wait ev_fork  ->  SSCAN(i): writes the sparse lists  ->  STAIL(i): reads them,
                                              writes the sparse half of `cands`  ->  wait ev_scan[p]  ->  DTAIL(i): reads
                                              lists[p] and thr[p], writes the dense half of `cands`  ->  record ev_tail[p]
                                              ->  FUSE(i): reads both halves of `cands`, writes the result block

Each stream is in order; an operation has a start and an end, and the scheduler may run anything of the OTHER stream in
between (a CUDA event wait refers to the record that preceded it in HOST order, which is how the model names them:
"DTAIL(i) waits for SCAN(i)", "SCAN(i) waits for DTAIL(i-2)").  Checked over thousands of random schedules: every read
sees exactly the data of its own search from start to end (no buffer overwritten while it is being read, none read before
it was written), and nothing deadlocks.  The candidate lists and threshold sets need TWO copies and the ev_tail wait; the
`cands` / result buffers get away with one because everything touching them is on the one in-order second stream.
Mutants (one copy of the lists, no ev_tail wait, the memset issued before the wait) must be caught, or the model would
prove nothing."""
import random

import pytest


class Model:
    def __init__(self, searches, copies=2, wait_tail=True, memset_before_wait=False, seed=0):
        self.S, self.copies, self.wait_tail, self.memset_before_wait = searches, copies, wait_tail, memset_before_wait
        self.rng = random.Random(seed)
        self.buf = {}                 # buffer name -> token (search index) of its current content
        self.reading = {}             # buffer name -> search index of the operation reading it right now
        self.done = set()             # finished operations: (name, i)
        self.errors = []
        self.streams = {"main": self._main_ops(), "side": self._side_ops()}
        self.pc = {"main": 0, "side": 0}
        self.open = {"main": None, "side": None}       # the operation a stream is in the middle of

    def _p(self, i):
        return i % self.copies

    def _main_ops(self):
        ops = []
        for i in range(self.S):
            wait = ("wait", ("DTAIL", i - 2)) if (self.wait_tail and i >= 2) else None
            memset = ("op", "MEMSET", i, [], [f"thr{self._p(i)}"])
            first = [memset, wait] if self.memset_before_wait else [wait, memset]
            ops += [o for o in first if o is not None]
            ops.append(("op", "FORK", i, [], []))
            ops.append(("op", "SCAN", i, [], [f"lists{self._p(i)}", f"thr{self._p(i)}"]))
        return ops

    def _side_ops(self):
        ops = []
        for i in range(self.S):
            ops.append(("wait", ("FORK", i)))
            ops.append(("op", "SSCAN", i, [], ["sp_lists"]))
            ops.append(("op", "STAIL", i, ["sp_lists"], ["cands_s"]))
            ops.append(("wait", ("SCAN", i)))
            ops.append(("op", "DTAIL", i, [f"lists{self._p(i)}", f"thr{self._p(i)}"], ["cands_d"]))
            ops.append(("op", "FUSE", i, ["cands_s", "cands_d"], ["out"]))
        return ops

    # ---- one scheduler step on one stream; returns False if the stream cannot move
    def step(self, st):
        if self.open[st] is not None:                          # finish the operation in flight
            _, name, i, reads, writes = self.open[st]
            for b in reads:
                if self.buf.get(b) != i:
                    self.errors.append(f"{name}({i}) finished reading {b} but it holds {self.buf.get(b)}")
                self.reading.pop(b, None)
            for b in writes:
                self.buf[b] = i
            self.done.add((name, i))
            self.open[st] = None
            self.pc[st] += 1
            return True
        if self.pc[st] >= len(self.streams[st]):
            return False
        op = self.streams[st][self.pc[st]]
        if op[0] == "wait":
            if op[1] not in self.done:
                return False
            self.pc[st] += 1
            return True
        _, name, i, reads, writes = op
        for b in reads:
            if self.buf.get(b) != i:
                self.errors.append(f"{name}({i}) starts reading {b} but it holds {self.buf.get(b)}")
            self.reading[b] = i
        for b in writes:
            if b in self.reading:
                self.errors.append(f"{name}({i}) overwrites {b} while search {self.reading[b]} is reading it")
            self.buf[b] = ("being written", i)
        self.open[st] = op
        return True

    def run(self):
        while True:
            order = ["main", "side"]
            self.rng.shuffle(order)
            # bias: let one stream run far ahead now and then (a slow tail, a slow scan)
            if self.rng.random() < 0.3:
                order = [order[0]] * 6 + order
            moved = False
            for st in order:
                if self.step(st):
                    moved = True
                    break
            if not moved:
                for st in ("main", "side"):
                    if self.step(st):
                        moved = True
                        break
            if not moved:
                break
        finished = all(self.pc[s] >= len(self.streams[s]) for s in self.streams)
        return finished, self.errors


@pytest.mark.parametrize("seed", range(40))
def test_pipelined_form_is_race_free_and_live(seed):
    finished, errors = Model(searches=12, seed=seed).run()
    assert finished, "deadlock"
    assert errors == []
    assert Model(searches=1, seed=seed).run() == (True, [])
    assert Model(searches=2, seed=seed).run() == (True, [])


def _caught(**mutant):
    return any(Model(searches=12, seed=s, **mutant).run()[1] for s in range(200))


def test_mutants_are_caught():
    assert _caught(copies=1), "one copy of the candidate lists must race (SCAN(i+1) against DTAIL(i))"
    assert _caught(wait_tail=False), "without the ev_tail wait SCAN(i+2) overwrites lists DTAIL(i) has not read"
    assert _caught(memset_before_wait=True), "clearing thr[p] before waiting for DTAIL(i-2) destroys its threshold"


def test_one_copy_is_enough_when_the_scan_waits_for_the_previous_tail():
    """The alternative the double buffering avoids: one copy + SCAN(i) waiting for DTAIL(i-1) is correct too, but then
    the scan of search i cannot start before the tail of search i-1 has run -- the serial form the pipelining removes."""
    class Serial(Model):
        def _main_ops(self):
            ops = []
            for i in range(self.S):
                if i >= 1:
                    ops.append(("wait", ("DTAIL", i - 1)))
                ops += [("op", "MEMSET", i, [], ["thr0"]), ("op", "FORK", i, [], []), ("op", "SCAN", i, [], ["lists0", "thr0"])]
            return ops
    for seed in range(40):
        assert Serial(searches=10, copies=1, seed=seed).run() == (True, [])
