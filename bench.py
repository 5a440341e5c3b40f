#!/usr/bin/env python
"""bench.py -- hybrid top-10 queries/sec & p50 latency over a 10M x 1024-d corpus on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]                   # the reference-shaped CPU path

One "step" = one batch of B queries through the whole hot path (dense scan + sparse scan + per-leg top-L + exact
re-score + RRF + top-k).  The corpus is FIXED at --rows (default 10M) and row-sharded over the N ranks
("scaling": "strong"); rank r generates its own row range on its GPU with the deterministic device generators.

  value       queries/s with every step's query batch already resident in HBM (one device slot per step, staged before
              the timed region): the K steps [legs -> candidate exchange -> fuse] are enqueued back to back on the
              launching stream and bracketed by one pair of CUDA events; MAX over ranks.  p50/p95 come from one event
              pair per step.
  e2e         the same metric through the host-buffer call a plugin makes (`b200rag_search` at N=1, else
              `ShardedSearcher.search`): fp32 host query vectors + sparse CSR in pinned/pageable host memory ->
              normalise -> H2D -> kernels -> (all-gather) -> D2H of ids/scores, wall clock, max over ranks.
  value_incl_copies
              the same K steps with NOTHING pre-staged: every step copies its query batch from pinned host memory, runs
              and reads its result back, all inside one CUDA event pair (SURVEY 8d's timing method).
  e2e_plugin  (N = 1) wall clock through the reference-facing plugin call itself: `B200Retriever.search` with an
              `EmbeddingResult` (Python lists) in and `RetrievalResult` objects out, the bench's shard adopted with
              `attach_prebuilt`.
  oracle_check
              OUTSIDE the timed region: the hits of the last two e2e steps are compared with the CPU oracle run over
              EVERY row of the corpus (each rank streams its shard's stored rows through oracle/oracle_c.c, the per-rank
              leg lists are gathered and fused by the oracle); mismatches are summed over ranks.
  roofline    the dominant kernel (dense_scan): algorithmic bytes (rows_on_this_rank * dim * 2 per launch) / its mean
              launch duration, measured with CUDA events inside the timed region, / the measured HBM copy peak.
  cpu_baseline / --impl reference
              the oracle's reference-shaped port (oracle.RefShapedIndex: fp32 sgemv + argsort, a Python two-pointer
              loop per document, dict RRF -- the algorithmic shape of qdrant-client local mode, which the reference
              runs with qdrant_in_memory=True) on a bounded row sample, scaled linearly to the full corpus.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-rag_b200")]

if any(x == "reference" or x.endswith("=reference") for x in sys.argv[1:]) and os.environ.get("RANK", "0") == "0":
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 ALONE (the others exit), so it
    # gets every host core its BLAS can use.  Must happen before numpy loads its BLAS.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        if os.environ.get(_v) == "1" and "TORCHELASTIC_RUN_ID" in os.environ:
            del os.environ[_v]

import numpy as np  # noqa: E402

SEED, QSEED = 1234, 2000
METRIC = "hybrid top-10 queries/sec, 10Mx1024-d corpus"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--top-k", type=int, default=10)
    ap.add_argument("--mode", default="hybrid", choices=["dense", "sparse", "hybrid"])
    ap.add_argument("--query-tokens", type=int, default=12)
    ap.add_argument("--cpu-sample-rows", type=int, default=20_000)
    ap.add_argument("--cpu-queries", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-dense-1m", action="store_true",
                    help="skip the CPU port's measured 1M-row dense-only figure (config 2 context, ~10 s)")
    ap.add_argument("--no-oracle-check", action="store_true", help="skip the full-corpus oracle check (profiling runs)")
    ap.add_argument("--no-compressed-leg", action="store_true",
                    help="skip the extra (non-headline) measurement of the opt-in 8-bit candidate scan")
    ap.add_argument("--compressed", action="store_true",
                    help="opt-in 8-bit candidate scan (b200rag_set_compression): NOT the headline configuration")
    return ap.parse_args()


def workload(a, world=None):
    """The `config` of BOTH arms (the reference arm reports on this arm's config): workload + how this arm runs it."""
    world = a.gpus if world is None else world
    per = -(-a.rows // world)
    per = min(a.rows, -(-per // 8192) * 8192)
    algo_gb = per * a.dim * 2 / 1e9
    return {"workload": f"{a.mode} dense(1024-d bf16 cosine)+sparse(Zipf BM25 impacts) RRF top-{a.top_k}, "
                        f"{a.rows} synthetic 256-token chunks, query batch {a.batch}",
            "corpus_rows": a.rows, "dim": a.dim, "batch": a.batch, "top_k": a.top_k, "search_type": a.mode,
            "query_terms": a.query_tokens, "vocab": 250_002, "rows_per_gpu": per,
            "parallelism": (f"row-sharded x{world}, one process per GPU, per-shard candidates exchanged over NVLink "
                            f"(CUDA-IPC peer windows, NCCL all-gather as fallback), merge/RRF kernel on every rank")
            if world > 1 else "one shard on one GPU",
            "l2": f"inputs larger than L2: {algo_gb:.2f} GB of corpus rows per GPU per step vs 126 MB L2; "
                  f"a different query every step",
            "timing": "one CUDA event pair around the K back-to-back steps on the launching stream (query "
                      "batches pre-staged in HBM), max over ranks; p50/p95 from per-step event pairs"}


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(a, steps, warmup, sample_rows):
    """Reference-shaped CPU path on a bounded sample; returns (qps_at_full_rows, detail dict)."""
    from b200rag import synth
    from oracle import fast, oracle
    n = min(sample_rows, a.rows)
    dense = synth.bf16_bits_to_f32(fast.synth_dense_bf16(SEED, 0, n, a.dim))      # fp32 unit rows, like qdrant-local
    thr = synth.zipf_thresholds(synth.VOCAB)
    idf, tff = synth.bm25_tables(a.rows)
    ip, tt, ww = fast.synth_sparse_csr(SEED, 0, n, thr, idf, tff, synth.VOCAB, 256, synth.TERM_PERM_MUL)
    ref = oracle.RefShapedIndex(dense, ip, tt, ww)
    nq = warmup + steps
    qf = synth.dense_queries_f32(QSEED, 0, nq, n, a.dim, corpus_seed=SEED)
    qi, qt, qw = synth.sparse_queries(QSEED, 0, nq, a.query_tokens, synth.VOCAB, thr)
    times = []
    for i in range(nq):
        sl = slice(qi[i], qi[i + 1])
        t0 = time.perf_counter()
        if a.mode == "hybrid":
            ref.hybrid(qf[i], qt[sl], qw[sl], None, a.top_k)
        elif a.mode == "dense":
            ref.dense_leg(qf[i], None, a.top_k)
        else:
            ref.sparse_leg(qt[sl], qw[sl], None, a.top_k)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_query_sample = float(np.mean(times))
    # once with the default BLAS threads (above), once pinned to one thread (a few queries are enough: the
    # single-threaded Python sparse loop dominates either way)
    one_thread_ms = None
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            t1 = []
            for i in range(warmup, min(nq, warmup + 4)):
                sl = slice(qi[i], qi[i + 1])
                t0 = time.perf_counter()
                if a.mode == "hybrid":
                    ref.hybrid(qf[i], qt[sl], qw[sl], None, a.top_k)
                elif a.mode == "dense":
                    ref.dense_leg(qf[i], None, a.top_k)
                else:
                    ref.sparse_leg(qt[sl], qw[sl], None, a.top_k)
                t1.append(time.perf_counter() - t0)
            one_thread_ms = float(np.mean(t1) * 1e3)
    except Exception:
        pass
    scale = a.rows / n
    qps = 1.0 / (per_query_sample * scale)
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        blas_threads = os.cpu_count()
    measured_10k = None
    try:      # BASELINE config 1, the reference's own CPU-runnable case, MEASURED (no scaling): hybrid top-5 over 10k rows
        n1 = min(10_000, n)
        ref1 = oracle.RefShapedIndex(dense[:n1], ip[:n1 + 1], tt[:ip[n1]], ww[:ip[n1]])
        t10 = []
        for i in range(min(nq, 10)):
            sl = slice(qi[i], qi[i + 1])
            t0 = time.perf_counter()
            ref1.hybrid(qf[i], qt[sl], qw[sl], None, 5)
            t10.append(time.perf_counter() - t0)
        measured_10k = {"rows": n1, "top_k": 5, "queries": len(t10), "ms_per_query": float(np.mean(t10) * 1e3),
                        "queries_per_s": float(1.0 / np.mean(t10)), "scaled": False}
    except Exception:
        pass
    measured_dense_1m = None
    if not a.no_cpu_dense_1m:
        try:  # SURVEY 8d "Tier B at 1 M dense-only for config 2 context": fp32 sgemv + argsort over 1M x 1024 rows, MEASURED
            n2 = 1_000_000
            d2 = synth.bf16_bits_to_f32(fast.synth_dense_bf16(SEED, 0, n2, a.dim))
            ref2 = oracle.RefShapedIndex(d2, np.zeros(n2 + 1, np.int64), np.zeros(0, np.uint32), np.zeros(0, np.float32))
            q2 = synth.dense_queries_f32(QSEED, 0, 8, n2, a.dim, corpus_seed=SEED)
            t2 = []
            for i in range(8):
                t0 = time.perf_counter()
                ref2.dense_leg(q2[i], None, 10)
                if i >= 2:
                    t2.append(time.perf_counter() - t0)
            measured_dense_1m = {"rows": n2, "top_k": 10, "search_type": "dense", "queries": len(t2),
                                 "ms_per_query": float(np.mean(t2) * 1e3), "queries_per_s": float(1.0 / np.mean(t2)),
                                 "scaled": False, "blas_threads": int(blas_threads)}
            del d2, ref2
        except Exception:
            measured_dense_1m = None
    detail = {"value": qps, "unit": "queries/s", "cores": int(blas_threads), "kind": "port",
              "measured_10k": measured_10k, "measured_dense_1m": measured_dense_1m, "measured_ms_per_query_on_sample": per_query_sample * 1e3,
              "sample_rows": n, "extrapolation_factor": scale,
              "sample": f"{len(times)} single hybrid queries over a {n}-row slice of the corpus "
                        f"({per_query_sample * 1e3:.1f} ms/query on the slice; BLAS sgemv uses {blas_threads} threads, "
                        f"the per-document sparse loop is single-threaded Python as in qdrant-client local mode), "
                        f"scaled x{scale:.0f} linearly to {a.rows} rows",
              "p50_ms_on_sample": float(np.median(times) * 1e3), "host_cpus": os.cpu_count(),
              "affinity_cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None,
              "ms_on_sample_blas_1_thread": one_thread_ms}
    return qps, detail, per_query_sample * 1e3


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    qps, detail, ms = cpu_reference(a, a.steps, a.warmup, a.cpu_sample_rows)
    # one step = one query batch on the bounded SAMPLE (ms_per_step is what was really timed); `value` is that rate
    # scaled linearly to the full corpus (cpu_baseline.extrapolation_factor), cpu_baseline.measured_10k is unscaled
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms * a.batch, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(a),
            "ms_per_step_is": "measured on the bounded sample (cpu_baseline.sample_rows rows), not scaled",
            "cpu_baseline": detail,
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_uuid):
        self.rows, self.proc, self.uuid = [], None, gpu_uuid

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            import atexit
            atexit.register(self._kill)          # never leave the sampler behind, whatever ends the run
        except Exception:
            self.proc = None

    def _kill(self):
        try:
            if self.proc is not None and self.proc.poll() is None:
                self.proc.kill()
        except Exception:
            pass

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in ln.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived in [t0, t1] (perf_counter; None = every sample)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons, pw = [], [], set(), []
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.06)]
        if not rows:                      # a run shorter than one sampling period: the nearest samples are all there is
            rows = [r for _, r in self.rows[-2:]]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- the CUDA arm
def build_shard(a, dev, lo, hi):
    import torch
    from b200rag import Shard, synth
    n = hi - lo
    sparse = a.mode != "dense"
    sh = Shard(dim=a.dim, device=dev.index, row_base=lo, reserve_rows=n, reserve_postings=int(n * 198) if sparse else 0)
    sh.set_stream(torch.cuda.current_stream().cuda_stream)
    thr = torch.from_numpy(synth.zipf_thresholds(synth.VOCAB).view(np.int64)).to(dev)
    idf_h, tff_h = synth.bm25_tables(a.rows)
    idf, tff = torch.from_numpy(idf_h).to(dev), torch.from_numpy(tff_h).to(dev)
    chunk = 1 << 20
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        bits = torch.empty((m, a.dim), dtype=torch.int16, device=dev)
        sh.synth_dense(SEED, lo + s, m, bits)
        if sparse:
            counts = torch.empty(m, dtype=torch.int64, device=dev)
            sh.synth_sparse(SEED, lo + s, m, 256, thr, idf, tff, synth.TERM_PERM_MUL, counts, None, None, None)
            indptr = torch.empty(m + 1, dtype=torch.int64, device=dev)
            sh.exclusive_scan_i64(counts, m, indptr)
            nnz = int(indptr[-1].item())
            terms = torch.empty(nnz, dtype=torch.int32, device=dev)
            w = torch.empty(nnz, dtype=torch.float32, device=dev)
            sh.synth_sparse(SEED, lo + s, m, 256, thr, idf, tff, synth.TERM_PERM_MUL, None, indptr, terms, w)
            sh.add_device(m, bits, indptr, terms, w, nnz)
            del counts, indptr, terms, w
        else:
            sh.add_device(m, bits)
        del bits
    if sparse:
        sh.build()
    if a.compressed:
        sh.set_compression(True)
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    return sh



def plugin_e2e(a, sh, qf, qi, qt, qw, steps, warmup, expect_last_ids, nsteps):
    """Wall clock through the reference-facing plugin call: B200Retriever.search(EmbeddingResult) -> [RetrievalResult]."""
    from b200rag.compat import EmbeddingResult, RetrievalConfig, SparseVector
    from b200rag.retriever import B200Retriever
    try:
        conf = RetrievalConfig(qdrant_in_memory=True, top_k=a.top_k, search_type=a.mode)
    except TypeError:
        conf = RetrievalConfig(top_k=a.top_k, search_type=a.mode)
    r = B200Retriever(conf, embedding_dim=a.dim)
    r.attach_prebuilt([sh], "bench", hybrid=a.mode != "dense")
    n = warmup + steps
    first = max(0, nsteps * a.batch - n)                # the LAST queries of the run (the e2e loop ended on them)
    embs = [EmbeddingResult(dense=[float(x) for x in qf[i]],
                            sparse=SparseVector(indices=[int(t) for t in qt[qi[i]:qi[i + 1]]],
                                                values=[float(w) for w in qw[qi[i]:qi[i + 1]]]))
            for i in range(first, first + n)]
    ts, res = [], None
    for i, e in enumerate(embs):
        t0 = time.perf_counter()
        res = r.search(e, collection_name="bench")
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    out = {"value": len(ts) / float(np.sum(ts)), "unit": "queries/s", "p50_ms": float(np.median(ts) * 1e3),
           "queries": len(ts), "api": "B200Retriever.search(EmbeddingResult lists) -> list[RetrievalResult]"}
    if expect_last_ids is not None:
        out["matches_c_abi_path"] = bool([int(x.chunk.text.split()[1]) for x in res] ==
                                         [int(v) for v in expect_last_ids[0][:len(res)]])
    r._shards = None                                    # the bench owns the shard
    return out


def oracle_check(a, sh, kept, step_arrays, world, rank, dev):
    """The e2e hits of the last two steps against the oracle over ALL rows (tests/fullscale.py), mismatches summed."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from b200rag import normalize_bf16
    from fullscale import LegJob, merge_topk, stream_oracle_legs
    from oracle import oracle
    t0 = time.time()
    L = 2 * a.top_k if a.mode == "hybrid" else a.top_k
    jobs, meta = [], []
    for slot, ids, sc, cnt in kept:
        f, ip, tt, ww = step_arrays(slot)
        qb = normalize_bf16(f)
        for b in range(min(a.batch, 2)):                # (bounded: the first two queries of a batch)
            jobs.append(LegJob(qb[b] if a.mode != "sparse" else None,
                               tt[ip[b]:ip[b + 1]] if a.mode != "dense" else None,
                               ww[ip[b]:ip[b + 1]] if a.mode != "dense" else None, L=L))
            meta.append((ids[b], sc[b], int(cnt[b])))
    stream_oracle_legs(sh, jobs)
    mine = [(j.dense, j.sparse) for j in jobs]
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    bad = 0
    for q, (ids, sc, cnt) in enumerate(meta):
        e = (np.zeros(0, np.int64), np.zeros(0, np.float32))
        dl, sl = e, e
        for r in range(world):
            dl = merge_topk(dl, allr[r][q][0], L)
            sl = merge_topk(sl, allr[r][q][1], L)
        if a.mode == "hybrid":
            ei, es = oracle.rrf_fuse([dl[0], sl[0]], a.top_k)
        elif a.mode == "dense":
            ei, es = dl[0], dl[1].astype(np.float64)
        else:
            ei, es = sl[0], sl[1].astype(np.float64)
        ok = cnt == len(ei) and np.array_equal(ids[:cnt], ei) and np.array_equal(sc[:cnt], es)
        bad += 0 if ok else 1
    tot = torch.tensor([bad], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    return {"queries": len(meta), "mismatches": int(tot.item()), "ranks": world, "seconds": round(time.time() - t0, 2),
            "what": "ids, counts and fp64 scores of the e2e results vs the CPU oracle scoring EVERY row of the corpus "
                    "(each rank streams its shard's stored rows through oracle_c.c; leg lists gathered, fused by the oracle); "
                    "mismatching queries summed over ranks"}


def run_b200(a):
    import torch
    import torch.distributed as dist
    from b200rag import _ffi, normalize_bf16, synth
    from b200rag.dist import ShardedSearcher, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if _ffi.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 device visible; this path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == a.gpus or world == 1, f"--gpus {a.gpus} but WORLD_SIZE={world}"

    sampler = None
    if rank == 0:
        sampler = ClockSampler(str(torch.cuda.get_device_properties(dev).uuid))
        if not sampler.uuid.startswith("GPU-"):
            sampler.uuid = "GPU-" + sampler.uuid
        sampler.start()       # nvidia-smi takes a while to deliver its first sample: it runs from here on, and only the
                              # samples that arrive during the timed loops are summarised (ClockSampler.stop)
    lo, hi = shard_bounds(a.rows, world, rank, align=8192)
    t_build = time.time()
    sh = build_shard(a, dev, lo, hi)
    t_build = time.time() - t_build
    ss = ShardedSearcher(sh, dev)
    B, K, W = a.batch, a.steps, a.warmup
    nsteps = W + K

    # host-side queries for every step (fp32 unit vectors + sparse CSR: what an embedder hands the plugin)
    thr_h = synth.zipf_thresholds(synth.VOCAB)
    qf = synth.dense_queries_f32(QSEED, 0, nsteps * B, a.rows, a.dim, corpus_seed=SEED)
    qi, qt, qw = synth.sparse_queries(QSEED, 0, nsteps * B, a.query_tokens, synth.VOCAB, thr_h)

    def step_arrays(i):
        s, e = i * B, (i + 1) * B
        return qf[s:e], (qi[s:e + 1] - qi[s]), qt[qi[s]:qi[e]], qw[qi[s]:qi[e]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ (1) device-resident timing -> value
    # Every step's query batch is staged into its own device slot BEFORE the timed region (inputs resident in HBM).
    # The K timed steps are then enqueued back to back -- legs -> candidate exchange -> fuse, no host synchronisation
    # in between -- and bracketed by ONE pair of CUDA events (plus one pair per step for p50/p95).
    sh.set_profiling(True)
    n_slots = min(nsteps, 2048)                       # (the library holds up to 4096 staged batches; queries repeat beyond)
    for i in range(n_slots):
        f, ip, tt, ww = step_arrays(i)
        ss.stage(a.mode, a.top_k, normalize_bf16(f), ip, tt, ww, slot=i)
    barrier()
    t_load0 = time.perf_counter()        # clocks are summarised from here (warm-up included) to the end of the e2e loop
    for i in range(W):
        ss.use_slot(i % n_slots)
        b = ss.run_staged()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e_begin.record()
    for i in range(W, nsteps):
        ss.use_slot(i % n_slots)
        ev[i - W][0].record()
        b = ss.run_staged()
        ss.record_end(ev[i - W][1])              # (exchange + fuse may run on their own stream)
    ss.record_end(e_end)
    barrier()
    wall_value = time.perf_counter() - wall0
    total_ms = float(e_begin.elapsed_time(e_end))
    step_ms = np.array([s.elapsed_time(e) for s, e in ev], dtype=np.float64)
    st = sh.stats()
    launches = (st["kernel_launches"] + (1 if world > 1 and not ss.p2p else 0)) * K     # (+ the NCCL kernel without P2P)
    dense_bytes, postings = st["dense_bytes"], st["sparse_postings"]
    dense_path, dense_passes = st["dense_path"], st["dense_passes"]
    dense_ms, sparse_ms, pre_ms, tail_ms = [], [], [], []
    for j in range(min(K, 64)):                       # event timings of the last timed steps (ring of 64 in the library)
        t = sh.stats_step(j)
        dense_ms.append(t["dense_scan_ms"]); sparse_ms.append(t["sparse_scan_ms"])
        pre_ms.append(t["pre_scan_ms"]); tail_ms.append(t["tail_ms"])
    ids_dev, sc_dev, cnt_dev, amb = ss.fetch(b)
    sh.set_profiling(False)

    # ------------------------------------------------------------------ (1b) the same steps with the copies inside
    # nothing pre-staged: every step copies its (already normalised) query batch from host memory, runs, and reads its
    # result back; ONE CUDA event pair around the K steps on the launching stream
    pre_q = [(normalize_bf16(step_arrays(i % n_slots)[0]),) + tuple(step_arrays(i % n_slots)[1:]) for i in range(W, nsteps)]
    barrier()
    c_begin, c_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c_begin.record()
    for qb_, ip_, tt_, ww_ in pre_q:
        ss.stage(a.mode, a.top_k, qb_, ip_, tt_, ww_)
        ss.fetch(ss.run_staged())
    c_end.record()
    barrier()
    copies_ms = float(c_begin.elapsed_time(c_end))
    del pre_q

    # ------------------------------------------------------------------ (2) end-to-end through the host-buffer call
    e2e_t = []
    kept = []                                        # (step, ids, scores, counts) of the last two e2e steps -> oracle_check
    h2d = B * a.dim * 2 + (B + 1) * 8 + 0 + B * 8
    barrier()
    for i in range(nsteps):
        f, ip, tt, ww = step_arrays(i % n_slots)
        if i == W:
            barrier()
        t0 = time.perf_counter()
        qb = normalize_bf16(f)
        if world == 1:
            r_ids, r_sc, r_cnt = sh.search(a.mode, a.top_k, qb, ip, tt, ww)
        else:
            r_ids, r_sc, r_cnt = ss.search(a.mode, a.top_k, qb, ip, tt, ww)
        if i >= W:
            e2e_t.append(time.perf_counter() - t0)
            h2d = max(h2d, B * a.dim * 2 + (B + 1) * 8 + len(tt) * 8 + B * 8)
        if i >= nsteps - 2:
            kept.append((i % n_slots, r_ids.copy(), r_sc.copy(), r_cnt.copy()))
    barrier()
    clocks = sampler.stop(t_load0, time.perf_counter()) if sampler else None   # both timed loops (device-resident + e2e)
    e2e_total = float(np.sum(e2e_t))
    d2h = B * a.top_k * 16 + (B + 1) * 4
    # the last e2e step and the last device-resident step used the same queries: results must agree
    same = bool(np.array_equal(r_ids, ids_dev) and np.array_equal(r_sc, sc_dev))

    # ------------------------------------------------------------------ (3) through the plugin call itself (N = 1)
    plugin = None
    if world == 1:
        try:
            plugin = plugin_e2e(a, sh, qf, qi, qt, qw, min(K, 100), W, r_ids if B == 1 else None, nsteps)
        except Exception as e:
            plugin = {"value": None, "error": repr(e)}

    # ------------------------------------------------------------------ (4) oracle check, outside every timed region
    try:
        ocheck = {"queries": 0, "mismatches": None, "skipped": "--no-oracle-check"} if a.no_oracle_check else \
            oracle_check(a, sh, kept, step_arrays, world, rank, dev)
    except Exception as e:
        ocheck = {"queries": 0, "mismatches": None, "error": repr(e)}

    # ------------------------------------------------------------------ (5) the opt-in 8-bit candidate scan, same steps
    # NOT the headline: the same staged steps once more with b200rag_set_compression on (candidates from an int8 copy of
    # the rows, exact re-score from the bf16 rows).  Both paths are exact, so the last step must return the same bytes.
    q8, q8_ms, q8_dense_ms = None, 0.0, 0.0
    if not a.compressed and not a.no_compressed_leg and B <= 2 and a.dim in (512, 1024) and a.mode != "sparse":
        ok = 1.0
        try:
            sh.set_compression(True)
        except Exception as e:
            ok, q8 = 0.0, {"value": None, "error": repr(e)}
        if world > 1:                                 # every rank or none: the exchange waits on all of them
            okt = torch.tensor([ok], dtype=torch.float64, device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            ok = float(okt.item())
        if ok > 0:
            try:
                sh.set_profiling(True)
                barrier()
                for i in range(W):
                    ss.use_slot(i % n_slots)
                    b8 = ss.run_staged()
                barrier()
                q_begin, q_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q_begin.record()
                for i in range(W, nsteps):
                    ss.use_slot(i % n_slots)
                    b8 = ss.run_staged()
                ss.record_end(q_end)
                barrier()
                q8_ms = float(q_begin.elapsed_time(q_end))
                st8 = sh.stats()
                d8 = [sh.stats_step(j)["dense_scan_ms"] for j in range(min(K, 64))]
                q8_dense_ms = float(np.mean(d8))
                ids8, sc8, cnt8, amb8 = ss.fetch(b8)
                sh.set_profiling(False)
                q8 = {"dense_path": int(st8["dense_path"]), "ambiguous_flags": int(amb8),
                      "matches_bf16_path": bool(np.array_equal(ids8, ids_dev) and np.array_equal(sc8, sc_dev)),
                      "scan_bytes_per_rank": int(st8["dense_bytes"])}
            except Exception as e:
                q8 = {"value": None, "error": repr(e)}
        try:
            sh.set_compression(False)
        except Exception:
            pass

    # ------------------------------------------------------------------ reduce over ranks (MAX)
    red = torch.tensor([total_ms, e2e_total, float(np.mean(dense_ms)), float(np.mean(sparse_ms)), wall_value,
                        copies_ms, float(np.mean(pre_ms)), float(np.mean(tail_ms)), q8_ms, q8_dense_ms],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    (total_ms, e2e_total, dense_k_ms, sparse_k_ms, wall_value, copies_ms, pre_max, tail_max, q8_ms,
     q8_dense_ms) = [float(x) for x in red.tolist()]

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, which = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(f"dense_scan_rows_{hi - lo}")
            except Exception:
                traffic = None
        rows_rank = hi - lo
        algo_gb = rows_rank * a.dim * 2 / 1e9
        if dense_path == 3:      # the 8-bit candidate scan reads (dim + 16) bytes per row
            algo_gb = rows_rank * (a.dim + 16) / 1e9
        # the dominant kernel: SIMT bulk scan (1-2 queries, one launch per corpus pass) or the tcgen05 GEMM (batches:
        # one launch per pass of <= 256 queries).  Per-launch figures: bytes = rows * dim * 2, flops = 2 * q * rows * dim.
        passes = max(int(dense_passes), 1)
        launch_ms = dense_k_ms / passes
        achieved = algo_gb / (launch_ms / 1e3) if dense_k_ms > 0 else 0.0
        roof = {"bound": "hbm", "kernel": "dense_scan_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": which}
        if dense_path == 2:
            tf_peak = 1413.9
            try:
                tf_peak = float(json.load(open(peaks_path)).get("bf16_tflops_sustained", tf_peak))
            except Exception:
                pass
            q_per_launch = B / passes
            tflops = 2.0 * q_per_launch * rows_rank * a.dim / (launch_ms / 1e3) / 1e12 if dense_k_ms > 0 else 0.0
            roof["kernel"] = "dense_gemm_kernel (tcgen05; CTA pairs when more than 128 queries share a pass)"
            roof["traffic"] = None
            roof["hbm_GBps"], roof["hbm_frac"] = achieved, achieved / peak
            roof["tensor_TFLOPs"], roof["tensor_frac"] = tflops, tflops / tf_peak
            if tflops / tf_peak > achieved / peak:      # past the ridge: the tensor pipe is the binding roof
                roof.update({"bound": "tensor", "achieved": tflops, "peak": tf_peak, "unit": "TFLOP/s",
                             "frac": tflops / tf_peak,
                             "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"})
        if dense_path == 3:
            roof["kernel"] = "dense_scan_q8_kernel (opt-in 8-bit candidate scan + exact re-score from the bf16 rows)"
            roof["traffic"] = None
            roof["bf16_equivalent_GBps"] = rows_rank * a.dim * 2 / 1e9 / (launch_ms / 1e3) if dense_k_ms > 0 else 0.0
        roof.update({"algorithmic_bytes_per_launch": int(algo_gb * 1e9), "launches_per_step": passes,
                     "kernel_ms": launch_ms, "kernel_share_of_step": dense_k_ms / (total_ms / K),
                     "sparse_scan_ms": sparse_k_ms, "sparse_postings_per_step": int(postings),
                     "sparse_algorithmic_GBps": postings * 6 / 1e9 / (sparse_k_ms / 1e3) if sparse_k_ms > 0 else 0.0,
                     "step_algorithmic_GBps": (dense_bytes + postings * 6) / 1e9 / (total_ms / K / 1e3),
                     "step_frac": (dense_bytes + postings * 6) / 1e9 / (total_ms / K / 1e3) / peak})
        qps = B * K / (total_ms / 1e3)
        line = {
            "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if dense_path != 3 else "bf16 (exact re-score) after an int8 candidate scan", "data": "synthetic",
            "config": workload(a, world) if not a.compressed else {**workload(a, world), "compressed_candidate_scan": True},
            "exchange": ("CUDA-IPC peer windows (NVLink stores + epoch flags)" + (", pipelined tail" if ss.pipeline else "")
                         if ss.p2p else "NCCL all-gather") if world > 1 else None,
            "p50_ms": float(np.median(step_ms)), "p95_ms": float(np.percentile(step_ms, 95)),
            "value_incl_copies": {"value": B * K / (copies_ms / 1e3), "unit": "queries/s", "ms_per_step": copies_ms / K,
                                  "note": "per step: H2D of the query batch + kernels + D2H of the result, one CUDA event "
                                          "pair around the K steps, max over ranks"},
            "wall_ms_per_step_incl_staging": wall_value / K * 1e3,
            "e2e": {"value": B * K / e2e_total, "unit": "queries/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "p50_ms": float(np.median(e2e_t) * 1e3),
                    "api": "b200rag_search (C ABI, host buffers)" if world == 1 else
                           "b200rag.dist.ShardedSearcher.search (host buffers; stage/legs/exchange/fuse through the C ABI, "
                           + ("candidates exchanged through CUDA-IPC peer windows)" if ss.p2p else "NCCL all-gather of the candidates)"),
                    "matches_device_path": same},
            "e2e_plugin": plugin,
            "oracle_check": ocheck,
            "gpu_launches": int(launches),
            "roofline": roof,
            "step_breakdown_ms": {"launch_to_scan": pre_max, "dense_scan": dense_k_ms,
                                  "scan_end_to_fuse_end": tail_max,
                                  "rank0": {"launch_to_scan": float(np.mean(pre_ms)), "dense_scan": float(np.mean(dense_ms)),
                                            "scan_end_to_fuse_end": float(np.mean(tail_ms))},
                                  "note": "CUDA events inside the timed steps, mean over the steps, MAX over ranks"},
            "clocks": clocks,
            "build_s": t_build, "ambiguous_flags": int(amb),
        }
        if q8 is not None:
            if q8_ms > 0 and "error" not in q8:
                q8_bytes = rows_rank * (a.dim + 16)
                q8.update({"value": B * K / (q8_ms / 1e3), "unit": "queries/s", "ms_per_step": q8_ms / K,
                           "dense_scan_ms": q8_dense_ms,
                           "scan_GBps": q8_bytes / 1e9 / (q8_dense_ms / 1e3) if q8_dense_ms > 0 else None,
                           "scan_frac_of_hbm_peak": q8_bytes / 1e9 / (q8_dense_ms / 1e3) / peak if q8_dense_ms > 0 else None,
                           "note": "opt-in (b200rag_set_compression): candidates from an int8 copy of the rows (dim + 16 bytes "
                                   "per row), ranked by a rigorous upper bound, then the same exact fp64 re-score from the "
                                   "bf16 rows as the headline path; device-resident timing like `value`, max over ranks; "
                                   "costs (dim + 16) bytes of HBM per row on top of the bf16 rows"})
            line["compressed_candidate_scan"] = q8
        if world == 1 and not a.no_cpu_baseline:
            try:
                _, detail, _ = cpu_reference(a, a.cpu_queries, 2, a.cpu_sample_rows)
                line["cpu_baseline"] = detail
            except Exception as e:  # the oracle is a checker, never a dependency of the measured path
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": 0, "kind": "port",
                                        "sample": f"failed: {e}"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was re-routed."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    a = parse()
    # NCCL / CUDA libraries print to fd 1 ("NCCL version ..."): keep the contract of exactly one JSON line on stdout
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
