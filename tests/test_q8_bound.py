"""Host restatement of the 8-bit candidate scan's arithmetic (audio-rag_b200/csrc/dense_q8.cu: quantize_rows_kernel, the
query quantisation in dense_scan_q8_kernel's prologue, the upper bound `ub` of its consumer loop), checked against the
property the exactness proof rests on:

    for EVERY row and query:   fl32(exact fp64 score)  <=  ub

The kernel ranks rows by `ub`; a row outside the retained candidates therefore has an exact score at or below the
weakest retained `ub`, and the leg tail's guard (eps = 1e-7) turns that into "the leg is exact" or into a retry on the
bf16 scan.  If `ub` could ever fall below the exact score, a true top-k row could be dropped silently, so the bound is
tested here on random and adversarial vectors with the same fp32 operations the device executes (IEEE fp32: the library
is built without fast-math).  The device sums the residuals and the query norms in a different order (per-lane chains +
butterfly); the 1.0001 factors on l1 / e2 / ||y|| cover any order (<= 1100 fp32 additions: relative error < 7e-5), and
the test checks the stored quantities against their fp64 values with that margin rather than bit for bit.  CPU only."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

F = np.float32
HALF = F(0.5005)          # kQ8Half


def bf16_round(x):
    """fp32 -> bf16 -> fp32 (round to nearest even), the rows and queries the engine stores."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16 << 16
    return u.astype(np.uint32).view(np.float32)


def quantize_rows(x):
    """quantize_rows_kernel: x [n, dim] bf16 values as fp32 -> int8 values, scale, l1, e2 (the row trailer)."""
    x = np.asarray(x, dtype=np.float32)
    m = np.abs(x).max(axis=1)
    scale = np.where(m > 0, m / F(127.0), F(1.0)).astype(np.float32)
    inv = (F(1.0) / scale).astype(np.float32)
    v = np.rint((x * inv[:, None]).astype(np.float32)).astype(np.int32)           # __float2int_rn
    v = np.clip(v, -127, 127)
    l1 = np.abs(v).sum(axis=1)
    # e = fmaf(-scale, v, x): one rounding of the exact value (exact in fp64: 24-bit x 8-bit product, aligned difference)
    e = (x.astype(np.float64) - scale.astype(np.float64)[:, None] * v).astype(np.float32)
    e2sum = (e.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32)
    tr_l1 = ((scale * l1.astype(np.float32)).astype(np.float32) * F(1.0001)).astype(np.float32)
    tr_e2 = (np.sqrt(e2sum).astype(np.float32) * F(1.0001)).astype(np.float32)
    return v, scale, tr_l1, tr_e2


def quantize_query(y):
    """prologue of dense_scan_q8_kernel: 14-bit integers, qs, ||y||_1 and ||y||_2 with their safety factors."""
    y = np.asarray(y, dtype=np.float32)
    m = np.abs(y).max()
    qs = F(m / F(8191.0)) if m > 0 else F(1.0)
    inv = F(F(1.0) / qs)
    yq = np.clip(np.rint((y * inv).astype(np.float32)).astype(np.int32), -8191, 8191)
    ql1 = F(np.abs(y).sum(dtype=np.float32) * F(1.0001))
    ql2 = F(np.sqrt((y * y).sum(dtype=np.float32)).astype(np.float32) * F(1.0001))
    return yq, qs, ql1, ql2


def upper_bounds(x, y):
    """ub of every row of x for the query y, operation by operation as in the consumer loop."""
    v, scale, tr_l1, tr_e2 = quantize_rows(x)
    yq, qs, ql1, ql2 = quantize_query(y)
    dot = v.astype(np.int64) @ yq.astype(np.int64)
    assert np.abs(dot).max(initial=0) < 2 ** 31, "the integer dot product must fit the dp2a accumulator"
    sh = ((scale * qs).astype(np.float32) * dot.astype(np.float32)).astype(np.float32)
    erow = np.minimum(((HALF * scale).astype(np.float32) * ql1).astype(np.float32), (tr_e2 * ql2).astype(np.float32))
    ub = (sh + erow).astype(np.float32)
    ub = (ub + ((HALF * qs) * tr_l1).astype(np.float32)).astype(np.float32)
    ub = (ub + (F(4e-7) * np.abs(sh)).astype(np.float32)).astype(np.float32)
    ub = (ub + F(1e-7)).astype(np.float32)
    return ub, dict(v=v, scale=scale, tr_l1=tr_l1, tr_e2=tr_e2, yq=yq, qs=qs, ql1=ql1, ql2=ql2, sh=sh, erow=erow)


def exact_scores(x, y):
    return (x.astype(np.float64) @ y.astype(np.float64)).astype(np.float32)       # what the canonical re-score emits


def unit(a):
    n = np.linalg.norm(a, axis=-1, keepdims=True)
    return np.where(n > 0, a / np.where(n > 0, n, 1), 0).astype(np.float32)


def rows_of_every_kind(rng, n, dim):
    g = unit(rng.standard_normal((n, dim)))
    spike = np.zeros((n, dim), np.float32)
    spike[np.arange(n), rng.integers(0, dim, n)] = 1.0
    spike = unit(spike + rng.standard_normal((n, dim)).astype(np.float32) * 0.02)     # one dominant component
    flat = unit(rng.choice([-1.0, 1.0], (n, dim)).astype(np.float32))                 # every |x_i| equal: scale * 127 = |x_i|
    halves = unit((rng.integers(-127, 128, (n, dim)) + 0.5).astype(np.float32))       # values near rounding boundaries
    sparse = unit(rng.standard_normal((n, dim)) * (rng.random((n, dim)) < 0.03))      # mostly zeros
    heavy = unit(rng.standard_t(1.5, (n, dim)))                                       # heavy tails: a few huge components
    tiny = (g * F(1e-30)).astype(np.float32)                                          # un-normalised, near the denormals
    zero = np.zeros((2, dim), np.float32)
    return np.concatenate([g, spike, flat, halves, sparse, heavy, tiny, zero])


@pytest.mark.parametrize("dim", [512, 1024])
def test_upper_bound_holds_on_random_and_adversarial_rows(dim):
    rng = np.random.default_rng(dim)
    x = bf16_round(rows_of_every_kind(rng, 300, dim))
    queries = [unit(rng.standard_normal(dim)), x[7].copy(), x[300 + 5].copy(), x[600 + 1].copy(),
               unit(np.eye(dim, dtype=np.float32)[3] + 0.01 * rng.standard_normal(dim)),
               unit(rng.standard_t(1.5, dim)), -x[11]]
    for qi, y in enumerate(queries):
        y = bf16_round(y)
        ub, d = upper_bounds(x, y)
        s = exact_scores(x, y)
        bad = np.nonzero(~(s <= ub))[0]
        assert bad.size == 0, f"dim {dim} query {qi}: ub below the exact score at rows {bad[:5]}: {s[bad[:5]]} > {ub[bad[:5]]}"
        # the stored error terms dominate their exact values whatever the summation order on the device
        e_true = x.astype(np.float64) - d["scale"].astype(np.float64)[:, None] * d["v"]
        # (rows near the denormals: squared residuals underflow in fp32 -- by at most sqrt(dim * 1.2e-38) = 3.5e-18 in e2,
        #  eleven orders below the 1e-7 absolute slack of ub)
        assert np.all(d["tr_e2"].astype(np.float64) >= np.linalg.norm(e_true, axis=1) * (1 + 2e-5) - 1e-17)
        assert np.all(d["tr_l1"].astype(np.float64) >= d["scale"].astype(np.float64) * np.abs(d["v"]).sum(axis=1) * (1 + 2e-5) - 1e-30)
        assert float(d["ql1"]) >= np.abs(y.astype(np.float64)).sum() * (1 + 2e-5)
        assert float(d["ql2"]) >= np.linalg.norm(y.astype(np.float64)) * (1 + 2e-5)
        # ... and the quantisation step is what the Hoelder term assumes: |x_i - scale q_i| <= 0.5005 * scale
        assert np.all(np.abs(e_true) <= 0.5005 * d["scale"].astype(np.float64)[:, None] + 1e-45)
        # not vacuous: the bound is a band of the expected width, not a blanket
        width = (ub.astype(np.float64) - s.astype(np.float64))
        assert np.all(width >= 0)
        g = slice(0, 300)                       # the Gaussian rows
        assert np.median(width[g]) < 0.02, "a band wider than ~0.6 sigma of the score distribution would keep thousands of rows"


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 2 ** 32 - 1), st.sampled_from([512, 1024]), st.floats(-3, 3), st.floats(0.0, 1.0))
def test_upper_bound_property(seed, dim, logmag, mix):
    """Unstructured search: rows of arbitrary magnitude (add() accepts any bf16 bits), a query anywhere between a random
    direction and one of the rows."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((24, dim)).astype(np.float32) * F(10.0 ** logmag)
    x[rng.integers(0, 24), rng.integers(0, dim)] *= F(50.0)
    x = bf16_round(x)
    y = bf16_round(unit(mix * unit(x[3]) + (1 - mix) * unit(rng.standard_normal(dim))))
    ub, _ = upper_bounds(x, y)
    s = exact_scores(x, y)
    assert np.all(s <= ub), (s - ub).max()


def test_measured_residual_bound_is_the_tighter_one_on_typical_rows():
    """DESIGN 4, "Opt-in": e2 * ||y||_2 is ~30 % below 0.5 * scale * ||y||_1 on Gaussian rows and queries -- which is
    why 236 spare candidates suffice where the first version needed 364."""
    rng = np.random.default_rng(3)
    x = bf16_round(unit(rng.standard_normal((500, 1024))))
    y = bf16_round(unit(rng.standard_normal(1024)))
    _, d = upper_bounds(x, y)
    holder = (HALF * d["scale"]).astype(np.float64) * float(d["ql1"])
    cs = d["tr_e2"].astype(np.float64) * float(d["ql2"])
    ratio = np.median(cs / holder)
    assert 0.6 < ratio < 0.8, ratio
    assert np.all(d["erow"].astype(np.float64) <= np.minimum(holder, cs) * (1 + 1e-6))
