"""Host model of leg_tail_kernel's survivor threshold (audio-rag_b200/csrc/select.cu, step 1), index for index:

  * NT threads; in round j thread t reads key  j * NT + ((t + 37 * j) mod NT)  (coalesced, rotated by 37 slots per round);
  * the rounds are dealt to C = ceil(kk / 32) classes, kk = ceil(Lc / #warps); per class every thread keeps the maximum
    of its keys, every warp publishes the kc-th largest of its 32 maxima (kc = ceil(kk / C) <= 32);
  * tau = the minimum of all published values.

Two properties: (1) VALIDITY -- at least Lc keys are >= tau, whatever the data (the merge may then cut at tau without
losing one of the best Lc); (2) STRENGTH on the shapes the scans produce -- n_lists SORTED lists of Lc keys -- including
the one that was pathological before the rotation: Lc dividing NT made thread t see rank t mod Lc of every list, the
maxima were stratified by rank and nearly every key survived (250 us of radix select on the GPU instead of 60 us).
CPU only; the kernel itself is covered by the GPU parity tests."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def tail_threshold(keys, Lc, NT, rotate=True):
    keys = np.asarray(keys, dtype=np.uint64)
    total = len(keys)
    NW = NT // 32
    kk = -(-Lc // NW)
    C = -(-kk // 32)
    kc = -(-kk // C)
    rounds = -(-total // NT)
    t = np.arange(NT)
    tau = None
    for cl in range(C):
        tmax = np.zeros(NT, dtype=np.uint64)
        for j in range(cl, rounds, C):
            idx = j * NT + ((t + (37 * j if rotate else 0)) % NT)
            ok = idx < total
            vals = np.zeros(NT, dtype=np.uint64)
            vals[ok] = keys[idx[ok]]
            tmax = np.maximum(tmax, vals)
        per_warp = np.sort(tmax.reshape(NW, 32), axis=1)[:, ::-1][:, kc - 1]      # kc-th largest of each warp's maxima
        m = per_warp.min()
        tau = m if tau is None else min(tau, m)
    return int(tau)


def sorted_lists(rng, n_lists, Lc, rows_per_list=70_000):
    """What a scan leaves behind: every CTA's best Lc of its own random share of the rows, sorted descending."""
    out = np.empty((n_lists, Lc), dtype=np.uint64)
    for i in range(n_lists):
        s = rng.standard_normal(rows_per_list).astype(np.float32)
        top = np.sort(s)[::-1][:Lc]
        # order-preserving key: positive floats compare like their bit patterns; shift everything positive first
        out[i] = ((top + np.float32(16.0)).view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.uint64(rng.integers(1, 2 ** 31))
    return out.reshape(-1)


@settings(max_examples=120, deadline=None)
@given(st.integers(0, 2 ** 32 - 1), st.sampled_from([256, 512, 1024]), st.integers(1, 768), st.integers(1, 40_000),
       st.booleans())
def test_threshold_is_valid_for_any_data(seed, NT, Lc, total, zeros):
    rng = np.random.default_rng(seed)
    keys = rng.integers(1, 2 ** 63, total, dtype=np.uint64)
    if zeros:
        keys[rng.random(total) < 0.5] = 0                  # empty list slots
    tau = tail_threshold(keys, Lc, NT)
    if tau == 0:
        return                                              # the kernel then keeps every non-empty slot
    assert int((keys >= np.uint64(tau)).sum()) >= min(Lc, int((keys > 0).sum()))


@pytest.mark.parametrize("n_lists,Lc,NT", [(148, 256, 1024), (148, 256, 512), (148, 384, 512), (148, 768, 512),
                                           (148, 768, 1024), (148, 300, 1024), (148, 26, 512)])
def test_threshold_is_strong_on_sorted_lists(n_lists, Lc, NT):
    rng = np.random.default_rng(Lc * 7 + NT)
    keys = sorted_lists(rng, n_lists, Lc, rows_per_list=20_000)
    tau = tail_threshold(keys, Lc, NT)
    survivors = int((keys >= np.uint64(tau)).sum())
    assert survivors >= Lc
    assert survivors <= 4096, f"{survivors} of {len(keys)} keys survive: the kernel would fall into its radix select"


def test_plain_stride_was_pathological_when_lc_divides_the_cta():
    """The regression this model documents: without the rotation, 148 sorted lists of 256 keys on 1 024 threads."""
    rng = np.random.default_rng(1)
    keys = sorted_lists(rng, 148, 256, rows_per_list=20_000)
    plain = int((keys >= np.uint64(tail_threshold(keys, 256, 1024, rotate=False))).sum())
    rotated = int((keys >= np.uint64(tail_threshold(keys, 256, 1024, rotate=True))).sum())
    assert plain > 4096 and plain > 5 * rotated, (plain, rotated)
