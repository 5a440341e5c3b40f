"""Host emulation of the sparse scan's fixed-point accumulation (audio-rag_b200/csrc/sparse.cu::sparse_scale and the
FFMA + IMAD increment), checked against the bound the kernel hands to the slack guard.

The GPU kernel selects candidates by APPROXIMATE scores and the guard must know how wrong they can be:
|approx - exact| <= q_eps = nterms / S + 2.4e-7 * sum|w_q| * max|w_d|.  This test restates the device arithmetic in numpy
(every step is exactly representable, so the emulation is bit-faithful) and verifies the bound, the count bits
("touched" semantics) and the no-overflow argument on random and adversarial inputs.  CPU only."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

MAGIC = np.float32(12582912.0)          # 1.5 * 2^23
MAGIC_BITS = 0x4B400000


def sparse_scale(qabs: float, w_absmax: float, nterms: int):
    """sparse.cu::sparse_scale: count bits and the power-of-two scale."""
    cb = nterms.bit_length()            # 32 - clz(nterms)
    bound = np.float32(qabs) * np.float32(w_absmax)
    e = 0
    if bound > 0 and bound < 3.0e38:
        bits = min(30 - cb, 22)
        e = bits - math.frexp(float(bound))[1] + 1 - 1      # ilogb(x) = frexp exponent - 1
        e = max(-60, min(60, e))
    return cb, np.float32(math.ldexp(1.0, e))


def accumulate(wq, wd, S, cb):
    """One document: every common term adds (round(w_q*S*w_d) << cb) + 1 to an int32 (two's complement wrap)."""
    acc = np.int64(0)
    for a, b in zip(wq, wd):
        wqs = np.float32(np.float32(a) * S)                             # exact: S is a power of two
        f = np.float32(np.float64(wqs) * np.float64(np.float32(b)) + np.float64(MAGIC))   # fmaf: ONE rounding
        c = int(f.view(np.uint32)) - MAGIC_BITS                         # rounded product, as an integer
        acc += (c << cb) + 1
    acc32 = int((int(acc) + 2 ** 31) % 2 ** 32 - 2 ** 31)               # what the 32-bit accumulator holds
    return acc32


def decode(acc32, S, cb):
    cnt = acc32 & ((1 << cb) - 1)
    v = acc32 >> cb                                                     # arithmetic shift
    return cnt, float(np.float32(np.float32(v) * (np.float32(1.0) / S)))


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 300), st.integers(0, 2 ** 32 - 1), st.floats(-6, 6), st.floats(-6, 6), st.booleans())
def test_fixed_point_error_is_within_the_guard_bound(nterms, seed, qmag, dmag, signed):
    rng = np.random.default_rng(seed)
    wq = (10.0 ** qmag * rng.uniform(0.01, 1.0, nterms)).astype(np.float32)
    wd_all = (10.0 ** dmag * rng.uniform(0.0, 1.0, nterms)).astype(np.float32)
    if signed:
        wq *= rng.choice([-1, 1], nterms).astype(np.float32)
        wd_all *= rng.choice([-1, 1], nterms).astype(np.float32)
    w_absmax = float(np.abs(wd_all).max()) if nterms else 0.0
    qabs = float(np.float32(np.abs(wq).astype(np.float32).sum(dtype=np.float32)))
    cb, S = sparse_scale(qabs, w_absmax, nterms)
    eps = nterms / float(S) + 2.4e-7 * qabs * w_absmax
    for _ in range(4):                                                  # documents sharing random subsets of the terms
        hit = rng.random(nterms) < rng.uniform(0.05, 1.0)
        if not hit.any():
            hit[rng.integers(nterms)] = True
        a32 = accumulate(wq[hit], wd_all[hit], S, cb)
        cnt, approx = decode(a32, S, cb)
        assert cnt == int(hit.sum()), "the low cb bits count the postings that touched the document"
        exact = float(np.sum(wq[hit].astype(np.float64) * wd_all[hit].astype(np.float64)))
        assert abs(approx - exact) <= eps * (1 + 1e-6), (approx, exact, eps, float(S), cb)


def test_scale_never_overflows_the_accumulator():
    """Worst case: every posting at +-max weight.  |sum| stays below 2^(31 - cb) so the count bits are never polluted."""
    for nterms in (1, 2, 15, 16, 255, 256, 4000):
        for wmax in (1e-4, 1.0, 37.5, 1e5):
            wq = np.full(nterms, 3.0, np.float32)
            cb, S = sparse_scale(float(np.abs(wq).sum(dtype=np.float32)), wmax, nterms)
            for sign in (1.0, -1.0):
                a32 = accumulate(wq, np.full(nterms, sign * wmax, np.float32), S, cb)
                cnt, approx = decode(a32, S, cb)
                assert cnt == nterms
                assert abs(approx - sign * 3.0 * wmax * nterms) <= nterms / float(S) + 2.4e-7 * 3.0 * nterms * wmax
                assert abs(a32 >> cb) < 2 ** (31 - cb)


def test_zero_weight_postings_are_touched_with_score_zero():
    cb, S = sparse_scale(2.0, 5.0, 3)
    a32 = accumulate(np.array([1.0, 1.0], np.float32), np.array([0.0, 0.0], np.float32), S, cb)
    assert decode(a32, S, cb) == (2, 0.0)
    assert decode(0, S, cb)[0] == 0            # an untouched accumulator is recognisable


@pytest.mark.parametrize("x", [0.5, 1.5, 2.5, -0.5, -1.5, 1234567.5, 4194303.49])
def test_magic_constant_rounds_to_nearest_even(x):
    f = np.float32(np.float64(x) + np.float64(MAGIC))
    assert int(f.view(np.uint32)) - MAGIC_BITS == int(np.rint(x))
