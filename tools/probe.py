"""Quick device-side timing probe (not the bench): dense / sparse / hybrid legs on a synthetic shard."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-rag_b200")]
from b200rag import Shard, normalize_bf16, synth  # noqa: E402


def build_shard(n, dim, sparse, dev, R=0, chunk=1 << 20, n_total=None, row0=0):
    n_total = n_total or n
    sh = Shard(dim=dim, device=dev.index, docs_per_block=R, row_base=row0, reserve_rows=n,
               reserve_postings=int(n * 200) if sparse else 0)
    sh.set_stream(torch.cuda.current_stream().cuda_stream)
    thr = torch.from_numpy(synth.zipf_thresholds(synth.VOCAB).view(np.int64)).to(dev)
    idf_h, tff_h = synth.bm25_tables(n_total)
    idf, tff = torch.from_numpy(idf_h).to(dev), torch.from_numpy(tff_h).to(dev)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        bits = torch.empty((m, dim), dtype=torch.int16, device=dev)
        sh.synth_dense(1234, row0 + s, m, bits)
        if sparse:
            counts = torch.empty(m, dtype=torch.int64, device=dev)
            sh.synth_sparse(1234, row0 + s, m, 256, thr, idf, tff, synth.TERM_PERM_MUL, counts, None, None, None)
            indptr = torch.empty(m + 1, dtype=torch.int64, device=dev)
            sh.exclusive_scan_i64(counts, m, indptr)
            nnz = int(indptr[-1].item())
            terms = torch.empty(nnz, dtype=torch.int32, device=dev)
            w = torch.empty(nnz, dtype=torch.float32, device=dev)
            sh.synth_sparse(1234, row0 + s, m, 256, thr, idf, tff, synth.TERM_PERM_MUL, None, indptr, terms, w)
            sh.add_device(m, bits, indptr, terms, w, nnz)
        else:
            sh.add_device(m, bits)
        del bits
    if sparse:
        t = time.time()
        sh.build()
        torch.cuda.synchronize()
        print(f"build inverted: {time.time() - t:.2f}s postings={sh.postings}", flush=True)
    return sh


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--modes", default="dense,sparse,hybrid")
    ap.add_argument("--batches", default="1,2")
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--R", type=int, default=0)
    ap.add_argument("--dense-path", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    modes = a.modes.split(",")
    sparse = any(m != "dense" for m in modes)
    t = time.time()
    sh = build_shard(a.rows, a.dim, sparse, dev, a.R)
    torch.cuda.synchronize()
    print(f"shard built in {time.time() - t:.1f}s rows={sh.count}", flush=True)
    sh.set_dense_path(a.dense_path)
    sh.set_profiling(True)
    for mode in modes:
        for B in [int(x) for x in a.batches.split(",")]:
            qf = synth.dense_queries_f32(2000, 0, B, a.rows, a.dim, corpus_seed=1234)
            ip, tt, ww = synth.sparse_queries(2000, 0, B)
            qb = normalize_bf16(qf)
            q, keep = sh.make_query(mode, a.topk, qb, ip, tt, ww)
            nlegs, L = Shard.legs_len(q)
            cands = torch.zeros((nlegs, B, L, 2), dtype=torch.int64, device=dev)
            amb = torch.zeros(1, dtype=torch.int32, device=dev)
            oi = torch.empty((B, a.topk), dtype=torch.int64, device=dev)
            osc = torch.empty((B, a.topk), dtype=torch.float64, device=dev)
            oc = torch.empty(B, dtype=torch.int32, device=dev)
            sh.stage(q, keep)
            for _ in range(3):
                sh.legs(cands, amb)
                sh.fuse(cands, 1, oi, osc, oc)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
            ev[0].record()
            for i in range(a.iters):
                sh.legs(cands, amb)
                sh.fuse(cands, 1, oi, osc, oc)
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters))
            st = sh.stats()
            p50 = ms[len(ms) // 2]
            gb = st["dense_bytes"] / 1e9
            print(f"mode={mode} B={B} p50={p50:.3f}ms min={ms[0]:.3f}ms launches={st['kernel_launches']} "
                  f"dense_GB={gb:.3f} dense_GB/s(p50)={gb / (p50 / 1e3):.0f} scan_ms={st['dense_scan_ms']:.3f} "
                  f"scan_GB/s={gb / max(st['dense_scan_ms'], 1e-9) * 1e3:.0f} TF={2.0 * B * a.rows * a.dim / max(st['dense_scan_ms'], 1e-9) / 1e9:.1f} sparse_ms={st['sparse_scan_ms']:.3f} amb={int(amb.item())} "
                  f"ids0={oi[0, :3].tolist()} cnt={oc[0].item()}", flush=True)
            # wall-clock through the host-buffer call
            t0 = time.perf_counter()
            for _ in range(a.iters):
                sh.search(mode, a.topk, qb, ip, tt, ww)
            print(f"   e2e search(): {(time.perf_counter() - t0) / a.iters * 1e3:.3f} ms/batch", flush=True)


if __name__ == "__main__":
    main()
