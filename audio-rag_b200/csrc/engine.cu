// C-ABI entry points of libb200rag.so (include/b200rag.h) and the host-side orchestration of one shard.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "common.cuh"
#include <cstdio>
#include <thread>

#include "engine.h"

namespace b200rag {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    if (e == cudaErrorMemoryAllocation) return B200RAG_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return B200RAG_ERR_NOGPU;
    return B200RAG_ERR_CUDA;
}

int DevBuf::ensure(size_t bytes, size_t keep_bytes, cudaStream_t st) {
    if (bytes <= cap && p != nullptr) return B200RAG_OK;
    size_t ncap = cap + cap / 2;
    if (ncap < bytes) ncap = bytes;
    ncap = (ncap + 255) & ~(size_t)255;
    if (ncap == 0) ncap = 256;
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess) {
        cudaGetLastError();
        // retry with the exact size before giving up
        ncap = (bytes + 255) & ~(size_t)255;
        e = cudaMalloc(&np, ncap);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    }
    if (p != nullptr) {
        if (keep_bytes > 0) {
            e = cudaMemcpyAsync(np, p, keep_bytes, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) { cudaFree(np); return cuda_fail(e, "cudaMemcpyAsync(grow)"); }
        }
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cudaFree(np); return cuda_fail(e, "cudaStreamSynchronize(grow)"); }
        cudaFree(p);
    }
    p = np;
    cap = ncap;
    return B200RAG_OK;
}
void DevBuf::release() {
    if (p != nullptr) cudaFree(p);
    p = nullptr;
    cap = 0;
}

__global__ void offset_copy_i64_kernel(int64_t* __restrict__ dst, const int64_t* __restrict__ src, int64_t n,
                                       int64_t offset) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + offset;
}

// device twin of b200rag_normalize_bf16: one thread per row, the host routine's fp64 operations in the host's order
__global__ void normalize_bf16_kernel(const float* __restrict__ x, int64_t n, int dim, uint16_t* __restrict__ out,
                                      int* __restrict__ bad) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float* row = x + r * (int64_t)dim;
    double ss = 0.0;
    bool finite = true;
    for (int k = 0; k < dim; ++k) {
        const double v = (double)row[k];
        finite = finite && isfinite(v);
        ss = __dadd_rn(ss, __dmul_rn(v, v));
    }
    if (!finite) { atomicExch(bad, 1); return; }
    double nrm = sqrt(ss);
    if (nrm == 0.0) nrm = 1.0;
    for (int k = 0; k < dim; ++k) {
        const float y = (float)__ddiv_rn((double)row[k], nrm);
        uint32_t u = __float_as_uint(y);
        u = u + 0x7FFFu + ((u >> 16) & 1u);
        out[r * (int64_t)dim + k] = (uint16_t)(u >> 16);
    }
}

// sparse query validation on the device: terms < vocab, strictly ascending inside each query
__global__ void validate_sparse_query_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ terms,
                                             int batch, uint32_t vocab, int* __restrict__ bad) {
    const int b = blockIdx.x;
    if (b >= batch) return;
    for (int64_t i = indptr[b] + threadIdx.x; i < indptr[b + 1]; i += blockDim.x) {
        if (terms[i] >= vocab) atomicExch(bad, 1);
        if (i > indptr[b] && terms[i] <= terms[i - 1]) atomicExch(bad, 2);
    }
}

static int use_device(const Shard* s) {
    B2_CUDA(cudaSetDevice(s->cfg.device));
    return B200RAG_OK;
}

// the shard's stream and, in pipelined mode, the side stream that carries the tails of the searches in flight
static int sync_all(Shard* s) {
    B2_CUDA(cudaStreamSynchronize(s->stream));
    if (s->pipeline && s->pipe_stream != nullptr) B2_CUDA(cudaStreamSynchronize(s->pipe_stream));
    return B200RAG_OK;
}

static int ensure_pinned(Shard* s, size_t bytes) {
    if (bytes <= s->h_pinned_cap) return B200RAG_OK;
    if (s->h_pinned != nullptr) { cudaStreamSynchronize(s->stream); cudaFreeHost(s->h_pinned); s->h_pinned = nullptr; }
    size_t ncap = std::max(bytes * 2, (size_t)1 << 16);
    B2_CUDA(cudaMallocHost(&s->h_pinned, ncap));
    s->h_pinned_cap = ncap;
    return B200RAG_OK;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// bring the 8-bit copy up to date with the rows stored (after appends: the new rows; after compaction / load: all)
static int q8_sync_rows(Shard* s) {
    if (!s->q8 || s->q8_rows == s->n_rows) return B200RAG_OK;
    const size_t rb = (size_t)s->dim + 16;
    if (s->q8_rows > s->n_rows) s->q8_rows = 0;
    B2_TRY(s->dense_q8.ensure((size_t)std::max<int64_t>(s->n_rows, 1) * rb, (size_t)s->q8_rows * rb, s->stream));
    B2_TRY(launch_quantize_rows(s, s->q8_rows, s->n_rows - s->q8_rows));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    s->q8_rows = s->n_rows;
    return B200RAG_OK;
}

// ids_host: global ids of the new rows (strictly increasing, above every id stored so far) or nullptr = row_base + local
static int append_rows(Shard* s, int64_t n, const uint16_t* dense, const int64_t* indptr, const uint32_t* terms,
                       const float* w, int64_t nnz, bool host, const int64_t* ids_host, int dense_on_device = -1) {
    cudaStream_t st = s->stream;
    B2_TRY(sync_all(s));               // (pipelined mode: tails still in flight read the buffers that grow below)
    const cudaMemcpyKind kind = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;   // CSR arrays
    const cudaMemcpyKind dkind = (dense_on_device < 0 ? !host : dense_on_device != 0) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t row_bytes = (size_t)s->dim * 2;
    int64_t new_last;
    if (ids_host != nullptr) {
        int64_t prev = s->last_id;
        for (int64_t i = 0; i < n; ++i) {
            if (ids_host[i] <= prev) { set_error("add: row ids must be strictly increasing over the shard's lifetime"); return B200RAG_ERR_INVALID; }
            prev = ids_host[i];
        }
        new_last = prev;
    } else {
        if (s->cfg.row_base + s->n_rows <= s->last_id) { set_error("add: implicit row id (row_base + local row) is not above the ids stored so far"); return B200RAG_ERR_INVALID; }
        new_last = s->cfg.row_base + s->n_rows + n - 1;
    }
    B2_TRY(s->row_ids.ensure((size_t)(s->n_rows + n) * 8, (size_t)s->n_rows * 8, st));
    if (ids_host != nullptr) B2_CUDA(cudaMemcpyAsync(s->row_ids.as<int64_t>() + s->n_rows, ids_host, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    else B2_TRY(launch_fill_row_ids(s, s->row_ids.as<int64_t>() + s->n_rows, s->cfg.row_base + s->n_rows, n));
    B2_TRY(s->dense.ensure((size_t)(s->n_rows + n) * row_bytes, (size_t)s->n_rows * row_bytes, st));
    B2_TRY(s->fwd_ptr.ensure((size_t)(s->n_rows + n + 1) * 8, (size_t)(s->n_rows + 1) * 8, st));
    B2_TRY(s->fwd_terms.ensure((size_t)(s->nnz + nnz + 1) * 4, (size_t)s->nnz * 4, st));
    B2_TRY(s->fwd_w.ensure((size_t)(s->nnz + nnz + 1) * 4, (size_t)s->nnz * 4, st));
    B2_CUDA(cudaMemcpyAsync(s->dense.as<uint8_t>() + (size_t)s->n_rows * row_bytes, dense, (size_t)n * row_bytes, dkind, st));
    int64_t* fp = s->fwd_ptr.as<int64_t>() + s->n_rows;  // fp[0] already holds s->nnz
    if (indptr == nullptr) {
        std::vector<int64_t> flat((size_t)n, s->nnz);
        B2_CUDA(cudaMemcpyAsync(fp + 1, flat.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaStreamSynchronize(st));
    } else if (host) {
        std::vector<int64_t> sh((size_t)n);
        for (int64_t i = 0; i < n; ++i) sh[(size_t)i] = indptr[i + 1] + s->nnz;
        B2_CUDA(cudaMemcpyAsync(fp + 1, sh.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        if (nnz > 0) {
            B2_CUDA(cudaMemcpyAsync(s->fwd_terms.as<uint32_t>() + s->nnz, terms, (size_t)nnz * 4, kind, st));
            B2_CUDA(cudaMemcpyAsync(s->fwd_w.as<float>() + s->nnz, w, (size_t)nnz * 4, kind, st));
        }
        B2_CUDA(cudaStreamSynchronize(st));
    } else {
        offset_copy_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fp + 1, indptr + 1, n, s->nnz);
        B2_CUDA(cudaGetLastError());
        if (nnz > 0) {
            B2_CUDA(cudaMemcpyAsync(s->fwd_terms.as<uint32_t>() + s->nnz, terms, (size_t)nnz * 4, kind, st));
            B2_CUDA(cudaMemcpyAsync(s->fwd_w.as<float>() + s->nnz, w, (size_t)nnz * 4, kind, st));
        }
        B2_CUDA(cudaStreamSynchronize(st));
    }
    s->n_rows += n;
    s->nnz += nnz;
    s->last_id = new_last;
    if (s->q8) B2_TRY(q8_sync_rows(s));
    return B200RAG_OK;
}

static int default_slack(int L) { return std::max(16, L / 2); }

// per-query side arrays of the sparse leg inside ws.q_eps: [B] u64 grid-wide threshold keys | [B] f32 error bounds of
// the approximate scores | [B] i32 grid-wide thresholds in fixed-point units
struct SparseAux { uint64_t* gkey; float* eps; int* gthr; };
static SparseAux sparse_aux(Shard* s, int B) {
    SparseAux a;
    a.gkey = s->ws.q_eps.as<uint64_t>();
    a.eps = reinterpret_cast<float*>(a.gkey + B);
    a.gthr = reinterpret_cast<int*>(a.eps + B);
    return a;
}

// Pipelined mode, classic-form fallbacks (tcgen05 batches, exhaustive legs, index rebuilds): the main stream first
// waits until the side stream has finished the earlier searches' tails (they read buffers the classic form reuses) ...
static int pipeline_drain(Shard* s) {
    B2_CUDA(cudaEventRecord(s->ev_join, s->pipe_stream));
    B2_CUDA(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
    return B200RAG_OK;
}
// ... and afterwards the side stream (where exchange and fuse are enqueued) waits for the legs the main stream ran
// exchange + fuse of a search run where its tails ran: on the second stream after pipelined legs, on `stream` otherwise
static bool tail_on_second_stream(const Shard* s) {
    return s->x_stream != nullptr && !s->pipeline_paused && !s->legs_classic;
}
static int pipeline_handover(Shard* s) {
    B2_CUDA(cudaEventRecord(s->ev_fork, s->stream));
    B2_CUDA(cudaStreamWaitEvent(s->pipe_stream, s->ev_fork, 0));
    return B200RAG_OK;
}

static int run_legs(Shard* s, b200rag_cand* cands, int32_t* ambiguous) {
    const b200rag_query& q = s->q;
    cudaStream_t st = s->stream;
    const int B = q.batch;
    const int nlegs = q.mode == B200RAG_HYBRID ? 2 : 1;
    const int L = q.mode == B200RAG_HYBRID ? 2 * q.top_k : q.top_k;
    int slack = s->slack > 0 ? s->slack : default_slack(L);
    int Lc = L + slack;
    if (Lc > 3 * B200RAG_MAX_TOPK) Lc = 3 * B200RAG_MAX_TOPK;

    // thr: per call parity [B grid-wide dense thresholds | B dense-scan tile counters] (two sets: in pipelined mode the
    // previous search's tail still reads its thresholds while this search's scan raises the other set) | postings counter
    B2_TRY(s->ws.thr.ensure((size_t)(4 * B + 1) * 8, 0, st));
    s->ws.post_count.p = s->ws.thr.as<uint64_t>() + 4 * B;
    B2_TRY(s->ws.exact.ensure((size_t)B * Lc * 8, 0, st));

    const bool want_dense = q.mode != B200RAG_SPARSE;
    const bool want_sparse = q.mode != B200RAG_DENSE;
    if (s->exhaustive) {
        // always-exact path: canonical score of every eligible row + full sort, no scan kernels, never ambiguous
        if (s->pipeline) B2_TRY(pipeline_drain(s));
        if (ambiguous != nullptr) B2_CUDA(cudaMemsetAsync(ambiguous, 0, 4, st));
        s->stats.exhaustive = 1;
        b200rag_cand* o = cands;
        if (want_dense) {
            if (s->n_rows == 0) B2_CUDA(cudaMemsetAsync(o, 0, (size_t)B * L * sizeof(b200rag_cand), st));
            else B2_TRY(launch_exhaustive_leg(s, false, B, L, q.has_threshold && q.mode == B200RAG_DENSE, q.score_threshold, o));
            o += (size_t)B * L;
        }
        if (want_sparse) {
            if (s->n_rows == 0 || s->nnz == 0 || s->staged_q_terms == 0) B2_CUDA(cudaMemsetAsync(o, 0, (size_t)B * L * sizeof(b200rag_cand), st));
            else B2_TRY(launch_exhaustive_leg(s, true, B, L, 0, 0.f, o));
        }
        s->legs_classic = s->pipeline && !s->pipeline_paused;
        return B200RAG_OK;
    }
    if (want_sparse && s->built_rows != s->n_rows) {
        if (s->pipeline) B2_TRY(pipeline_drain(s));
        B2_TRY(build_inverted(s));
    }

    // Hybrid: the sparse leg (scan + fused tail) runs on the side stream while the dense scan streams the corpus.
    // The dense scan is issued FIRST so its persistent CTAs (1 per SM, 3-stage ring = 105 KB) are resident, and the
    // sparse CTAs (39 KB, 64 registers) co-reside with them.  Measured on B200: overlapped beats back-to-back at
    // 1.25M, 10M and 12.5M rows, top-10 and top-100 (10M: 3.26 vs 3.36 ms per search).
    const bool use_gemm_path = s->dense_path == 2 || (s->dense_path == 0 && B > 2);

    // 8-bit candidate scan (opt-in, 1-2 queries per pass, first attempt only)
    const bool use_q8 = s->q8 && want_dense && !use_gemm_path && s->slack == 0 && s->n_rows > 0 && s->q8_rows == s->n_rows &&
                        leg_tail_fits(dense_scan_nlists(s), 3 * B200RAG_MAX_TOPK);
    // (the error band of the upper bounds holds a multiple of L rows: a few dozen at top-10, several hundred at top-100)
    const int Lc_q8 = std::min(L + std::max(s->q8_slack, 3 * L), 3 * B200RAG_MAX_TOPK);

    // ---- pipelined form: the SIMT scan alone on the main stream, everything else on the side stream -------------
    const int nl_scan = dense_scan_nlists(s);
    // (only while the tails are light, i.e. up to 64 candidates per leg: the tails of search i run beside the dense scan
    //  of search i+1, whose tiles are split statically over the SMs, so a heavy tail -- top-100: 300 candidates to
    //  re-score -- holds back the SM it shares and with it the whole scan: 12.5M rows, top-100: 4.37 ms per step pipelined,
    //  4.19 ms in the classic form.  Claiming tiles dynamically instead costs more than it saves: B200RAG_SCAN_DYNAMIC.)
    const bool piped = s->pipeline && !s->pipeline_paused && s->pipe_stream != nullptr && s->fused_tail && !use_gemm_path && s->n_rows > 0 &&
                       (!use_q8 || (s->q8_pipeline && leg_tail_fits(nl_scan, Lc_q8))) && Lc <= 64 && leg_tail_fits(nl_scan, Lc) &&
                       (!want_sparse || leg_tail_fits(sparse_scan_nlists(s, B, Lc), Lc));
    if (piped) {
        cudaStream_t sd = s->pipe_stream;
        const int par = (int)(s->legs_calls & 1);                 // (legs_calls was incremented by the caller)
        DevBuf& lists = par ? s->ws.lists_b : s->ws.lists_a;      // the dense tail two calls ago has released this set
        b200rag_cand* out_dense = cands;
        b200rag_cand* out_sparse = cands + (want_dense ? (size_t)B * L : 0);
        const bool sparse_live = want_sparse && s->nnz > 0 && s->staged_q_terms > 0;
        int nlists = 0;
        const int Lc_d = use_q8 ? Lc_q8 : Lc;                     // candidates per list of the dense leg
        if (want_dense) {
            B2_TRY(lists.ensure((size_t)B * nl_scan * Lc_d * 8, 0, st));
            if (s->ev_tail_rec[par]) B2_CUDA(cudaStreamWaitEvent(st, s->ev_tail[par], 0));
            B2_CUDA(cudaMemsetAsync(s->ws.thr.as<uint64_t>() + (size_t)par * 2 * B, 0, (size_t)2 * B * 8, st));
        }
        B2_CUDA(cudaEventRecord(s->ev_fork, st));                  // side-stream work of this search starts after this point
        B2_CUDA(cudaStreamWaitEvent(sd, s->ev_fork, 0));
        if (ambiguous != nullptr) B2_CUDA(cudaMemsetAsync(ambiguous, 0, 4, sd));
        if (want_dense) {
            s->dense_stage_cap = s->dense_stage_cap_env;
            s->thr_par = par;
            const int rc = use_q8 ? launch_dense_scan_q8(s, B, Lc_d, lists.as<uint64_t>(), &nlists)
                                  : launch_dense_scan(s, B, Lc, lists.as<uint64_t>(), &nlists);
            s->thr_par = 0;
            s->dense_stage_cap = 0;
            if (rc != B200RAG_OK) return rc;
            B2_CUDA(cudaEventRecord(s->ev_scan[par], st));
        }
        s->stream = sd;                                            // the launchers below enqueue on s->stream
        s->tail_beside_scan = true;                                // (this search's scan, then the next one's)
        int rc = B200RAG_OK;
        if (want_sparse) {
            if (!sparse_live) {
                cudaError_t e = cudaMemsetAsync(out_sparse, 0, (size_t)B * L * sizeof(b200rag_cand), sd);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemsetAsync(sparse leg)");
            } else {
                const int sp_lists = sparse_scan_nlists(s, B, Lc);
                rc = s->ws.lists_c.ensure((size_t)B * sp_lists * Lc * 8, 0, sd);
                if (rc == B200RAG_OK) rc = s->ws.q_eps.ensure((size_t)B * 16, 0, sd);
                SparseAux aux = sparse_aux(s, B);
                if (rc == B200RAG_OK) {
                    cudaError_t e = cudaMemsetAsync(aux.gkey, 0, (size_t)B * 8, sd);
                    if (e == cudaSuccess) e = cudaMemsetAsync(aux.gthr, 0x80, (size_t)B * 4, sd);
                    if (e == cudaSuccess) e = cudaMemsetAsync(s->ws.post_count.p, 0, 8, sd);
                    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemsetAsync(sparse thresholds)");
                }
                if (rc == B200RAG_OK) rc = launch_sparse_scan(s, B, Lc, s->ws.lists_c.as<uint64_t>(), aux.eps, aux.gthr, aux.gkey);
                if (rc == B200RAG_OK) rc = launch_leg_tail(s, true, B, sp_lists, Lc, L, s->ws.lists_c.as<uint64_t>(), 1e-12f, 2e-6f,
                                                           aux.eps, 0, 0.f, out_sparse, ambiguous, aux.gkey);
            }
        }
        if (rc == B200RAG_OK && want_dense) {
            cudaError_t e = cudaStreamWaitEvent(sd, s->ev_scan[par], 0);
            if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamWaitEvent(scan)");
            const int dthr = q.has_threshold && q.mode == B200RAG_DENSE;
            if (rc == B200RAG_OK) rc = launch_leg_tail(s, false, B, nlists, Lc_d, L, lists.as<uint64_t>(), use_q8 ? 1e-7f : 6.5e-5f,
                                                       0.f, nullptr, dthr, q.score_threshold, out_dense, ambiguous,
                                                       s->ws.thr.as<uint64_t>() + (size_t)par * 2 * B);
            if (rc == B200RAG_OK) {
                e = cudaEventRecord(s->ev_tail[par], sd);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventRecord(tail)"); else s->ev_tail_rec[par] = true;
            }
        }
        s->stream = st;
        s->tail_beside_scan = false;
        s->legs_classic = false;
        return rc;
    }
    if (s->pipeline) B2_TRY(pipeline_drain(s));     // classic form below: the side stream's earlier searches must be through
    if (ambiguous != nullptr) B2_CUDA(cudaMemsetAsync(ambiguous, 0, 4, st));   // callers need not pre-zero the counter
    B2_CUDA(cudaMemsetAsync(s->ws.thr.p, 0, (size_t)(4 * B + 1) * 8, st));

    // Batched hybrid (tcgen05 path): the list epilogues need most of the register file and all of shared memory, so
    // nothing could co-reside.  The FILTER epilogue keeps no per-query state (96 registers); with one pipeline stage
    // less it leaves room for a sparse CTA per SM, and the two legs overlap like in the single-query case.
    const bool overlap_base = want_dense && want_sparse && s->overlap_legs && s->side_stream != nullptr && s->n_rows > 0 &&
                              (s->overlap_max_rows <= 0 || s->n_rows <= s->overlap_max_rows);
    // (only while the dense leg is HBM-bound, i.e. up to 128 queries per pass: measured on B200, 12.5M rows, B = 1024
    //  top-100: overlapped 80.6 ms, back to back 77.0 ms -- a tensor-bound GEMM and the sparse scan fight for issue slots)
    const bool overlap_batched = overlap_base && use_gemm_path && s->overlap_gemm && s->gemm_filter && s->slack == 0 &&
                                 s->nnz > 0 && s->staged_q_terms > 0 && (B <= 128 || s->overlap_gemm == 2);
    const bool overlap = overlap_base && (!use_gemm_path || overlap_batched);
    b200rag_cand* out_dense = cands;
    b200rag_cand* out_sparse = cands + (want_dense ? (size_t)B * L : 0);
    if (overlap) {
        B2_CUDA(cudaEventRecord(s->ev_fork, st));
        B2_CUDA(cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0));
    }
    if (want_dense) {
        b200rag_cand* out = out_dense;
        if (s->n_rows == 0) {
            B2_CUDA(cudaMemsetAsync(out, 0, (size_t)B * L * sizeof(b200rag_cand), st));
        } else {
            const int nl_max = dense_scan_nlists(s);
            B2_TRY(s->ws.lists_a.ensure((size_t)B * nl_max * Lc * 8, 0, st));
            B2_TRY(s->ws.lists_b.ensure((size_t)B * nl_max * Lc * 8, 0, st));
            int nlists = 0;
            const bool use_gemm = s->dense_path == 2 || (s->dense_path == 0 && B > 2);
            if (use_q8) {
                // 8-bit candidate scan: half the bytes, upper-bound keys, a wider candidate set (the rows inside the
                // quantisation error band), then the usual exact re-score from the bf16 rows.  A retry (slack != 0)
                // never comes here: it takes the bf16 scan.
                B2_TRY(s->ws.lists_a.ensure((size_t)B * nl_max * Lc_q8 * 8, 0, st));
                B2_TRY(s->ws.exact.ensure((size_t)B * Lc_q8 * 8, 0, st));
                s->dense_stage_cap = s->dense_stage_cap_env;
                const int rcq = launch_dense_scan_q8(s, B, Lc_q8, s->ws.lists_a.as<uint64_t>(), &nlists);
                s->dense_stage_cap = 0;
                if (rcq != B200RAG_OK) return rcq;
                const int dthr_q = q.has_threshold && q.mode == B200RAG_DENSE;
                B2_TRY(launch_leg_tail(s, false, B, nlists, Lc_q8, L, s->ws.lists_a.as<uint64_t>(), 1e-7f, 0.f, nullptr, dthr_q,
                                       q.score_threshold, out, ambiguous, s->ws.thr.as<uint64_t>()));
            } else {
            s->dense_stage_cap = s->dense_stage_cap_env;   // (3 x 32 KB leaves room for a co-resident sparse CTA too)
            uint64_t* approx = nullptr;
            // tcgen05 path with a top-k too large for register lists: sample + filter (128 queries per pass whatever
            // the top-k).  A retry (slack widened after an ambiguous result) takes the robust list path instead.
            const bool filtered = use_gemm && (Lc > 64 || overlap_batched) && s->slack == 0 && s->gemm_filter;
            s->gemm_smem_reserve = overlap_batched ? (size_t)64 * 1024 : 0;
            if (filtered) {
                B2_TRY(launch_dense_gemm_filtered(s, B, Lc, s->ws.lists_a.as<uint64_t>(), s->ws.lists_b.as<uint64_t>(),
                                                  s->ws.lists_a.as<uint64_t>() + (size_t)B * nl_max * Lc - (size_t)B * Lc,
                                                  ambiguous));
                approx = s->ws.lists_a.as<uint64_t>() + (size_t)B * nl_max * Lc - (size_t)B * Lc;
            } else {
                if (use_gemm) B2_TRY(launch_dense_gemm(s, B, Lc, s->ws.lists_a.as<uint64_t>(), &nlists, nullptr));
                else B2_TRY(launch_dense_scan(s, B, Lc, s->ws.lists_a.as<uint64_t>(), &nlists));
                if (!(s->fused_tail && leg_tail_fits(nlists, Lc)))
                    B2_TRY(launch_merge_tree(s, B, nlists, Lc, s->ws.lists_a.as<uint64_t>(), s->ws.lists_b.as<uint64_t>(), &approx));
            }
            s->dense_stage_cap = 0;
            s->gemm_smem_reserve = 0;
            const int dthr = q.has_threshold && q.mode == B200RAG_DENSE;
            if (approx == nullptr) {
                B2_TRY(launch_leg_tail(s, false, B, nlists, Lc, L, s->ws.lists_a.as<uint64_t>(), 6.5e-5f, 0.f, nullptr, dthr,
                                       q.score_threshold, out, ambiguous, s->ws.thr.as<uint64_t>()));
            } else {
                B2_TRY(launch_rescore_dense(s, B, Lc, approx, s->ws.exact.as<uint64_t>()));
                B2_TRY(launch_finalize_leg(s, B, Lc, L, approx, s->ws.exact.as<uint64_t>(), 6.5e-5f, 0.f, nullptr, dthr,
                                           q.score_threshold, out, ambiguous));
            }
            }   // !use_q8
        }
    }
    if (want_sparse) {
        b200rag_cand* out = out_sparse;
        cudaStream_t sst = overlap ? s->side_stream : st;
        if (s->n_rows == 0 || s->nnz == 0 || s->staged_q_terms == 0) {
            B2_CUDA(cudaMemsetAsync(out, 0, (size_t)B * L * sizeof(b200rag_cand), sst));
        } else {
            const int sp_lists = sparse_scan_nlists(s, B, Lc);
            const size_t need = (size_t)B * sp_lists * Lc * 8;
            B2_TRY(s->ws.lists_c.ensure(need, 0, st));
            B2_TRY(s->ws.lists_d.ensure(need, 0, st));
            B2_TRY(s->ws.exact2.ensure((size_t)B * Lc * 8, 0, st));
            B2_TRY(s->ws.q_eps.ensure((size_t)B * 16, 0, st));
            const SparseAux aux = sparse_aux(s, B);
            B2_CUDA(cudaMemsetAsync(aux.gkey, 0, (size_t)B * 8, sst));
            B2_CUDA(cudaMemsetAsync(aux.gthr, 0x80, (size_t)B * 4, sst));   // thresholds: 0x80808080 < any score
            s->stream = sst;                                   // the launchers below enqueue on s->stream
            s->tail_beside_scan = overlap && want_dense;       // (on the side stream, next to the dense scan's CTAs)
            int rc = launch_sparse_scan(s, B, Lc, s->ws.lists_c.as<uint64_t>(), aux.eps, aux.gthr, aux.gkey);
            if (s->fused_tail && leg_tail_fits(sp_lists, Lc)) {
                if (rc == B200RAG_OK) rc = launch_leg_tail(s, true, B, sp_lists, Lc, L, s->ws.lists_c.as<uint64_t>(), 1e-12f, 2e-6f,
                                                           aux.eps, 0, 0.f, out, ambiguous, aux.gkey);
            } else {
                uint64_t* approx = nullptr;
                if (rc == B200RAG_OK) rc = launch_merge_tree(s, B, sp_lists, Lc, s->ws.lists_c.as<uint64_t>(), s->ws.lists_d.as<uint64_t>(), &approx);
                if (rc == B200RAG_OK) rc = launch_rescore_sparse(s, B, Lc, approx, s->ws.exact2.as<uint64_t>());
                if (rc == B200RAG_OK) rc = launch_finalize_leg(s, B, Lc, L, approx, s->ws.exact2.as<uint64_t>(), 1e-12f, 2e-6f, aux.eps, 0, 0.f, out, ambiguous);
            }
            s->stream = st;
            s->tail_beside_scan = false;
            if (rc != B200RAG_OK) return rc;
        }
    }
    if (overlap) {
        B2_CUDA(cudaEventRecord(s->ev_join, s->side_stream));
        B2_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
    }
    s->legs_classic = s->pipeline && !s->pipeline_paused;      // exchange + fuse follow on this stream, then the hand-over
    (void)nlegs;
    return B200RAG_OK;
}

// host-side validation of a doc-major CSR handed to add (indptr from 0 and monotone, terms ascending/unique/in range)
static int validate_csr_host(const Shard* s, int64_t n, const int64_t* indptr, const uint32_t* terms, const float* w,
                             int64_t* nnz_out) {
    *nnz_out = 0;
    if (indptr == nullptr) return B200RAG_OK;
    if (indptr[0] != 0) { set_error("add: sparse indptr must start at 0"); return B200RAG_ERR_INVALID; }
    const int64_t nnz = indptr[n];
    if (nnz > 0 && (terms == nullptr || w == nullptr)) { set_error("add: sparse terms/weights missing"); return B200RAG_ERR_INVALID; }
    for (int64_t d = 0; d < n; ++d) {
        if (indptr[d + 1] < indptr[d]) { set_error("add: sparse indptr not monotone"); return B200RAG_ERR_INVALID; }
        for (int64_t i = indptr[d]; i < indptr[d + 1]; ++i) {
            if (terms[i] >= (uint32_t)s->vocab) { set_error("add: sparse index out of vocabulary range"); return B200RAG_ERR_INVALID; }
            if (i > indptr[d] && terms[i] <= terms[i - 1]) { set_error("add: sparse indices must be ascending and unique per row"); return B200RAG_ERR_INVALID; }
            if (!isfinite(w[i])) { set_error("add: non-finite sparse weight"); return B200RAG_ERR_INVALID; }
        }
    }
    *nnz_out = nnz;
    return B200RAG_OK;
}

}  // namespace b200rag

using namespace b200rag;

static void p2p_release(Shard* s);

extern "C" {

const char* b200rag_version(void) { return "b200rag 0.1 (sm_100a)"; }
const char* b200rag_last_error(void) { return g_err.c_str(); }

int b200rag_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

// rows [r0, r1): the reference arithmetic of rule R2 (fp64 sum of squares in index order, fp64 division, RNE to bf16)
static bool normalize_rows(const float* x, int64_t r0, int64_t r1, int32_t dim, uint16_t* out) {
    for (int64_t r = r0; r < r1; ++r) {
        const float* row = x + r * (int64_t)dim;
        double ss = 0.0;
        for (int k = 0; k < dim; ++k) {
            const double v = (double)row[k];
            if (!isfinite(v)) return false;
            ss = ss + v * v;
        }
        double nrm = sqrt(ss);
        if (nrm == 0.0) nrm = 1.0;
        for (int k = 0; k < dim; ++k) {
            const float y = (float)((double)row[k] / nrm);
            uint32_t u;
            memcpy(&u, &y, 4);
            u = u + 0x7FFFu + ((u >> 16) & 1u);
            out[r * (int64_t)dim + k] = (uint16_t)(u >> 16);
        }
    }
    return true;
}

int b200rag_normalize_bf16(const float* x, int64_t n, int32_t dim, uint16_t* out) {
    if (x == nullptr || out == nullptr || n < 0 || dim <= 0) { set_error("normalize: bad argument"); return B200RAG_ERR_INVALID; }
    // rows are independent: batches are split over the OpenMP team (a pooled team: spawning std::threads per call cost
    // ~0.5 ms for a 256-query batch, a tenth of the search it precedes; single-threaded the batch takes ~0.8 ms)
    int nthreads = (int)std::min<int64_t>(16, n / 16);
    bool ok = true;
    if (nthreads <= 1) {
        ok = normalize_rows(x, 0, n, dim, out);
    } else {
        int good = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(&& : good)
        for (int t = 0; t < nthreads; ++t) {
            const int64_t r0 = n * t / nthreads, r1 = n * (t + 1) / nthreads;
            good = good && (normalize_rows(x, r0, r1, dim, out) ? 1 : 0);
        }
        ok = good != 0;
    }
    if (!ok) { set_error("normalize: non-finite input"); return B200RAG_ERR_INVALID; }
    return B200RAG_OK;
}

int b200rag_shard_create(const b200rag_config* cfg, b200rag_shard** out) {
    if (cfg == nullptr || out == nullptr) { set_error("shard_create: null argument"); return B200RAG_ERR_INVALID; }
    *out = nullptr;
    if (cfg->dim <= 0 || cfg->dim % 256 != 0 || cfg->dim > 1024) { set_error("shard_create: dim must be a multiple of 256, <= 1024"); return B200RAG_ERR_INVALID; }
    if (cfg->vocab <= 0) { set_error("shard_create: vocab must be positive"); return B200RAG_ERR_INVALID; }
    int R = cfg->docs_per_block == 0 ? 8192 : cfg->docs_per_block;
    if (R < 1024 || R > 16384 || (R & (R - 1)) != 0) { set_error("shard_create: docs_per_block must be a power of two in [1024, 16384]"); return B200RAG_ERR_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: this library has no CPU fallback");
        return B200RAG_ERR_NOGPU;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("shard_create: bad device ordinal"); return B200RAG_ERR_INVALID; }
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        set_error(std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) + ", this library is built for sm_100a only");
        return B200RAG_ERR_NOGPU;
    }
    B2_CUDA(cudaSetDevice(cfg->device));
    Shard* s = new Shard();
    s->cfg = *cfg;
    s->dim = cfg->dim;
    s->vocab = cfg->vocab;
    s->R = R;
    s->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete s; return cuda_fail(e, "cudaStreamCreate"); }
    s->stream = s->own_stream;
    if (cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        s->side_stream = nullptr;   // no overlap, still correct
    }
    if (const char* e = getenv("B200RAG_OVERLAP")) { s->overlap_legs = atoi(e) != 0; s->overlap_force = atoi(e) == 2; }
    if (const char* e = getenv("B200RAG_OVERLAP_MAX_ROWS")) s->overlap_max_rows = atoll(e);
    if (const char* e = getenv("B200RAG_DENSE_STAGES")) s->dense_stage_cap_env = atoi(e);
    if (const char* e = getenv("B200RAG_TILE_INTERLEAVE")) s->tile_interleave = atoi(e);
    if (const char* e = getenv("B200RAG_SCAN_SHARED")) s->scan_shared = atoi(e);
    if (const char* e = getenv("B200RAG_GEMM_FILTER")) s->gemm_filter = atoi(e) != 0;
    if (const char* e = getenv("B200RAG_GEMM_PAIRS")) s->gemm_pairs = atoi(e) != 0;
    if (const char* e = getenv("B200RAG_GEMM_L2_PREFETCH")) s->gemm_l2_prefetch = atoi(e) != 0;
    if (const char* e = getenv("B200RAG_GEMM_STAGES")) s->gemm_stage_cap = atoi(e);
    if (const char* e = getenv("B200RAG_FUSED_TAIL")) s->fused_tail = atoi(e) != 0;
    if (const char* e = getenv("B200RAG_OVERLAP_GEMM")) s->overlap_gemm = atoi(e);
    if (const char* e = getenv("B200RAG_SPARSE_THREADS")) s->sparse_threads = atoi(e);
    if (const char* e = getenv("B200RAG_SPARSE_BPC")) s->sparse_bpc = atoi(e);
    if (const char* e = getenv("B200RAG_Q8_SLACK")) { const int v = atoi(e); if (v > 0) s->q8_slack = v; }
    if (const char* e = getenv("B200RAG_Q8_PIPELINE")) s->q8_pipeline = atoi(e) != 0;
    if (const char* e = getenv("B200RAG_SCAN_DYNAMIC")) s->scan_dynamic = atoi(e) != 0;
    if (const char* e = getenv("B200RAG_SCAN_CTAS")) s->scan_ctas = atoi(e);
    if (const char* e = getenv("B200RAG_EXACT_FALLBACK")) s->exact_fallback = atoi(e) != 0;
    if (const char* e = getenv("B200RAG_P2P_TIMEOUT_MS")) { const long long ms = atoll(e); if (ms > 0) s->x_timeout_cycles = ms * 2000000ll; }
    if (const char* e = getenv("B200RAG_BULK_SPLIT")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) s->bulk_split = v; }
    int rc = s->fwd_ptr.ensure((size_t)(std::max<int64_t>(cfg->reserve_rows, 1024) + 1) * 8, 0, s->stream);
    if (rc == B200RAG_OK) {
        e = cudaMemsetAsync(s->fwd_ptr.p, 0, 8, s->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemsetAsync");
    }
    if (rc == B200RAG_OK && cfg->reserve_rows > 0) rc = s->dense.ensure((size_t)cfg->reserve_rows * s->dim * 2, 0, s->stream);
    if (rc == B200RAG_OK && cfg->reserve_postings > 0) {
        rc = s->fwd_terms.ensure((size_t)cfg->reserve_postings * 4, 0, s->stream);
        if (rc == B200RAG_OK) rc = s->fwd_w.ensure((size_t)cfg->reserve_postings * 4, 0, s->stream);
    }
    if (rc != B200RAG_OK) { b200rag_shard_destroy((b200rag_shard*)s); return rc; }
    *out = (b200rag_shard*)s;
    return B200RAG_OK;
}

void b200rag_shard_destroy(b200rag_shard* sp) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) return;
    cudaSetDevice(s->cfg.device);
    cudaStreamSynchronize(s->stream);
    if (s->side_stream) cudaStreamSynchronize(s->side_stream);
    if (s->pipeline && s->pipe_stream) cudaStreamSynchronize(s->pipe_stream);
    p2p_release(s);
    s->dense.release(); s->dense_q8.release(); s->row_ids.release(); s->fwd_ptr.release(); s->fwd_terms.release(); s->fwd_w.release();
    s->ws.ex_keys.release(); s->ws.ex_sorted.release(); s->ws.ex_temp.release();
    s->dir.release(); s->blk_base.release(); s->post_doc.release(); s->post_w.release();
    for (auto& kv : s->masks) kv.second.release();
    for (auto& sl : s->slots) sl.buf.release();
    s->ws.flag.release();
    s->ws.thr.release(); s->ws.lists_a.release(); s->ws.lists_b.release();
    s->ws.exact.release(); s->ws.pool.release(); s->ws.cands.release(); s->ws.out.release();
    s->ws.lists_c.release(); s->ws.lists_d.release(); s->ws.exact2.release(); s->ws.q_eps.release();
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    for (int r = 0; r < kProfileRing; ++r)
        for (int i = 0; i < 6; ++i)
            if (s->ev_ring[r][i]) cudaEventDestroy(s->ev_ring[r][i]);
    for (int i = 0; i < 2; ++i) {
        if (s->ev_scan[i]) cudaEventDestroy(s->ev_scan[i]);
        if (s->ev_tail[i]) cudaEventDestroy(s->ev_tail[i]);
    }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->side_stream) cudaStreamDestroy(s->side_stream);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
}

int b200rag_set_stream(b200rag_shard* sp, void* stream) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    B2_TRY(sync_all(s));
    s->stream = (cudaStream_t)stream;  // NULL == the legacy default stream
    return B200RAG_OK;
}

int b200rag_set_slack(b200rag_shard* sp, int32_t slack) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || slack < 0) { set_error("set_slack: bad argument"); return B200RAG_ERR_INVALID; }
    s->slack = slack;
    return B200RAG_OK;
}

int b200rag_set_compression(b200rag_shard* sp, int32_t on) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    B2_TRY(sync_all(s));
    if (!on) {
        s->q8 = false;
        s->q8_rows = 0;
        s->dense_q8.release();
        return B200RAG_OK;
    }
    if (!dense_q8_supported(s)) { set_error("set_compression: the 8-bit scan needs dim 512 or 1024"); return B200RAG_ERR_INVALID; }
    s->q8 = true;
    return q8_sync_rows(s);
}

int b200rag_set_exhaustive(b200rag_shard* sp, int32_t on) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    s->exhaustive = on != 0;
    return B200RAG_OK;
}

int b200rag_set_exact_fallback(b200rag_shard* sp, int32_t on) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    s->exact_fallback = on != 0;
    return B200RAG_OK;
}

int b200rag_set_pipeline(b200rag_shard* sp, int32_t on, void* stream) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    cudaStream_t second = stream != nullptr ? (cudaStream_t)stream : s->side_stream;
    if (on && second == nullptr) { set_error("set_pipeline: no second stream available"); return B200RAG_ERR_STATE; }
    if (on && second == s->stream) { set_error("set_pipeline: the second stream must differ from the shard's stream"); return B200RAG_ERR_INVALID; }
    B2_TRY(sync_all(s));
    if (s->side_stream != nullptr) B2_CUDA(cudaStreamSynchronize(s->side_stream));
    if (on && s->ev_scan[0] == nullptr)
        for (int i = 0; i < 2; ++i) {
            B2_CUDA(cudaEventCreateWithFlags(&s->ev_scan[i], cudaEventDisableTiming));
            B2_CUDA(cudaEventCreateWithFlags(&s->ev_tail[i], cudaEventDisableTiming));
        }
    s->ev_tail_rec[0] = s->ev_tail_rec[1] = false;
    s->pipeline = on != 0 ? 1 : 0;
    s->pipe_stream = s->pipeline ? second : nullptr;
    s->x_stream = s->pipe_stream;                            // exchange + fuse follow the tails
    return B200RAG_OK;
}

int b200rag_pipeline_pause(b200rag_shard* sp, int32_t on) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    s->pipeline_paused = on != 0;
    return B200RAG_OK;
}

void* b200rag_result_stream(const b200rag_shard* sp) {
    const Shard* s = (const Shard*)sp;
    if (s == nullptr) return nullptr;
    return (void*)(s->x_stream != nullptr && !s->pipeline_paused ? s->x_stream : s->stream);
}

int b200rag_set_dense_path(b200rag_shard* sp, int32_t path) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || path < 0 || path > 2) { set_error("set_dense_path: bad argument"); return B200RAG_ERR_INVALID; }
    s->dense_path = path;
    return B200RAG_OK;
}

int b200rag_debug_dense_scores(b200rag_shard* sp, float* out_scores_dev) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || out_scores_dev == nullptr) { set_error("debug_dense_scores: null argument"); return B200RAG_ERR_INVALID; }
    if (!s->staged || s->q.mode == B200RAG_SPARSE) { set_error("debug_dense_scores: stage a dense/hybrid batch first"); return B200RAG_ERR_STATE; }
    if (s->n_rows == 0) return B200RAG_OK;
    B2_TRY(use_device(s));
    const int B = s->q.batch, Lc = 16;
    B2_TRY(s->ws.thr.ensure((size_t)(4 * B + 1) * 8, 0, s->stream));
    s->ws.post_count.p = s->ws.thr.as<uint64_t>() + 4 * B;
    B2_CUDA(cudaMemsetAsync(s->ws.thr.p, 0, (size_t)(4 * B + 1) * 8, s->stream));
    B2_TRY(s->ws.lists_a.ensure((size_t)B * dense_gemm_nlists(s) * Lc * 8, 0, s->stream));
    int nlists = 0;
    B2_TRY(launch_dense_gemm(s, B, Lc, s->ws.lists_a.as<uint64_t>(), &nlists, out_scores_dev));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    return B200RAG_OK;
}

int b200rag_sync(b200rag_shard* sp) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    return sync_all(s);
}

int b200rag_add(b200rag_shard* sp, int64_t n, const uint16_t* dense, const int64_t* indptr, const uint32_t* terms,
                const float* w) {
    return b200rag_add_ids(sp, n, dense, indptr, terms, w, nullptr);
}

int b200rag_add_ids(b200rag_shard* sp, int64_t n, const uint16_t* dense, const int64_t* indptr, const uint32_t* terms,
                    const float* w, const int64_t* ids) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || n < 0 || (n > 0 && dense == nullptr)) { set_error("add: bad argument"); return B200RAG_ERR_INVALID; }
    if (n == 0) return B200RAG_OK;
    if (s->n_rows + n > 0xFFFFFFF0ll) { set_error("add: shard row limit (2^32) exceeded"); return B200RAG_ERR_INVALID; }
    int64_t nnz = 0;
    B2_TRY(validate_csr_host(s, n, indptr, terms, w, &nnz));
    B2_TRY(use_device(s));
    return append_rows(s, n, dense, indptr, terms, w, nnz, true, ids);
}

int b200rag_add_f32(b200rag_shard* sp, int64_t n, const float* dense_f32, const int64_t* indptr, const uint32_t* terms,
                    const float* w, const int64_t* ids) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || n < 0 || (n > 0 && dense_f32 == nullptr)) { set_error("add_f32: bad argument"); return B200RAG_ERR_INVALID; }
    if (n == 0) return B200RAG_OK;
    if (s->n_rows + n > 0xFFFFFFF0ll) { set_error("add_f32: shard row limit (2^32) exceeded"); return B200RAG_ERR_INVALID; }
    int64_t nnz = 0;
    B2_TRY(validate_csr_host(s, n, indptr, terms, w, &nnz));
    B2_TRY(use_device(s));
    // raw fp32 rows -> device -> unit bf16 rows (the device twin of b200rag_normalize_bf16), in slices of <= 64k rows
    DevBuf xf, xb;
    const int64_t slice = 65536;
    int rc = xf.ensure((size_t)std::min(n, slice) * s->dim * 4, 0, s->stream);
    if (rc == B200RAG_OK) rc = xb.ensure((size_t)n * s->dim * 2, 0, s->stream);
    for (int64_t r0 = 0; r0 < n && rc == B200RAG_OK; r0 += slice) {
        const int64_t m = std::min(slice, n - r0);
        cudaError_t e = cudaMemcpyAsync(xf.p, dense_f32 + (size_t)r0 * s->dim, (size_t)m * s->dim * 4, cudaMemcpyHostToDevice, s->stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(add_f32)"); break; }
        rc = b200rag_normalize_bf16_device(sp, xf.as<float>(), m, xb.as<uint16_t>() + (size_t)r0 * s->dim);
    }
    if (rc == B200RAG_OK) rc = append_rows(s, n, xb.as<uint16_t>(), indptr, terms, w, nnz, true, ids, 1);
    xf.release(); xb.release();
    return rc;
}

int b200rag_add_device(b200rag_shard* sp, int64_t n, const uint16_t* dense, const int64_t* indptr,
                       const uint32_t* terms, const float* w, int64_t nnz) {
    return b200rag_add_device_ids(sp, n, dense, indptr, terms, w, nnz, nullptr);
}

int b200rag_add_device_ids(b200rag_shard* sp, int64_t n, const uint16_t* dense, const int64_t* indptr,
                           const uint32_t* terms, const float* w, int64_t nnz, const int64_t* ids) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || n < 0 || nnz < 0 || (n > 0 && dense == nullptr)) { set_error("add_device: bad argument"); return B200RAG_ERR_INVALID; }
    if (n == 0) return B200RAG_OK;
    if (s->n_rows + n > 0xFFFFFFF0ll) { set_error("add_device: shard row limit (2^32) exceeded"); return B200RAG_ERR_INVALID; }
    if (indptr == nullptr) nnz = 0;
    B2_TRY(use_device(s));
    return append_rows(s, n, dense, indptr, terms, w, nnz, false, ids);
}

int b200rag_compact(b200rag_shard* sp, const uint32_t* keep_words, int64_t n_rows_mask) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || keep_words == nullptr) { set_error("compact: null argument"); return B200RAG_ERR_INVALID; }
    if (n_rows_mask != s->n_rows) { set_error("compact: the keep mask must cover exactly the shard's rows"); return B200RAG_ERR_INVALID; }
    if (s->n_rows == 0) return B200RAG_OK;
    B2_TRY(use_device(s));
    B2_TRY(sync_all(s));
    const size_t words = (size_t)((s->n_rows + 31) / 32);
    DevBuf km;
    B2_TRY(km.ensure(words * 4, 0, s->stream));
    cudaError_t e = cudaMemcpyAsync(km.p, keep_words, words * 4, cudaMemcpyHostToDevice, s->stream);
    if (e != cudaSuccess) { km.release(); return cuda_fail(e, "cudaMemcpyAsync(keep mask)"); }
    const int rc = compact_rows(s, km.as<uint32_t>());
    km.release();
    if (rc != B200RAG_OK) return rc;
    // local rows moved: masks and staged batches (which hold local bit positions / raw mask pointers) are void
    for (auto& kv : s->masks) kv.second.release();
    s->masks.clear(); s->mask_rows.clear();
    s->staged = false;
    for (auto& sl : s->slots) sl.staged = false;
    s->q8_rows = 0;                      // local rows moved: the 8-bit copy is rebuilt from the bf16 rows
    return q8_sync_rows(s);
}

int b200rag_build(b200rag_shard* sp) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    B2_TRY(sync_all(s));
    return build_inverted(s);
}

int64_t b200rag_count(const b200rag_shard* sp) { return sp ? ((const Shard*)sp)->n_rows : 0; }
int64_t b200rag_postings(const b200rag_shard* sp) { return sp ? ((const Shard*)sp)->nnz : 0; }

int b200rag_clear(b200rag_shard* sp) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    B2_TRY(sync_all(s));
    s->n_rows = 0; s->nnz = 0; s->built_rows = 0; s->n_blocks = 0; s->inv_nnz = 0;
    s->w_absmax = 0.f; s->wmax_nnz = 0;
    s->last_id = INT64_MIN;
    s->q8_rows = 0;
    B2_CUDA(cudaMemsetAsync(s->fwd_ptr.p, 0, 8, s->stream));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    s->h_blk_base.clear();
    s->staged = false;
    for (auto& sl : s->slots) sl.staged = false;
    for (auto& kv : s->masks) kv.second.release();
    s->masks.clear(); s->mask_rows.clear();
    return B200RAG_OK;
}

int b200rag_read_dense(b200rag_shard* sp, int64_t row, int64_t n, uint16_t* out) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || row < 0 || n < 0 || row + n > s->n_rows || out == nullptr) { set_error("read_dense: bad range"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    B2_CUDA(cudaMemcpyAsync(out, s->dense.as<uint16_t>() + (size_t)row * s->dim, (size_t)n * s->dim * 2, cudaMemcpyDeviceToHost, s->stream));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    return B200RAG_OK;
}

int b200rag_read_sparse(b200rag_shard* sp, int64_t row, int64_t n, int64_t* indptr_out, uint32_t* terms_out,
                        float* weights_out, int64_t cap_nnz) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || row < 0 || n < 0 || row + n > s->n_rows || indptr_out == nullptr) { set_error("read_sparse: bad range"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    B2_CUDA(cudaMemcpyAsync(indptr_out, s->fwd_ptr.as<int64_t>() + row, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s->stream));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    const int64_t base = indptr_out[0];
    for (int64_t i = 0; i <= n; ++i) indptr_out[i] -= base;
    const int64_t nnz = indptr_out[n];
    if (terms_out == nullptr || weights_out == nullptr) return B200RAG_OK;
    if (cap_nnz < nnz) { set_error("read_sparse: output arrays are smaller than the rows' postings"); return B200RAG_ERR_INVALID; }
    if (nnz > 0) {
        B2_CUDA(cudaMemcpyAsync(terms_out, s->fwd_terms.as<uint32_t>() + base, (size_t)nnz * 4, cudaMemcpyDeviceToHost, s->stream));
        B2_CUDA(cudaMemcpyAsync(weights_out, s->fwd_w.as<float>() + base, (size_t)nnz * 4, cudaMemcpyDeviceToHost, s->stream));
        B2_CUDA(cudaStreamSynchronize(s->stream));
    }
    return B200RAG_OK;
}

int b200rag_read_row_ids(b200rag_shard* sp, int64_t row, int64_t n, int64_t* out) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || row < 0 || n < 0 || row + n > s->n_rows || out == nullptr) { set_error("read_row_ids: bad range"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    if (n > 0) {
        B2_CUDA(cudaMemcpyAsync(out, s->row_ids.as<int64_t>() + row, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
        B2_CUDA(cudaStreamSynchronize(s->stream));
    }
    return B200RAG_OK;
}

static int mask_store(Shard* s, int32_t id, const uint32_t* words, int64_t n_rows, bool host) {
    if (id < 0 || words == nullptr || n_rows < 0) { set_error("mask_set: bad argument"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    // sized to whole inverted-index blocks so kernels may read any word of a touched block
    const int64_t cover = std::max<int64_t>(((std::max(n_rows, s->n_rows) + 32767) / 32768) * 32768, 32768);
    const size_t words_total = (size_t)(cover / 32);
    const size_t words_in = (size_t)((n_rows + 31) / 32);
    for (auto& sl : s->slots) sl.staged = false;     // staged batches hold raw mask pointers
    s->staged = false;
    B2_TRY(sync_all(s));                             // no search in flight may still read the mask being replaced
    DevBuf& b = s->masks[id];
    B2_TRY(b.ensure(words_total * 4, 0, s->stream));
    B2_CUDA(cudaMemsetAsync(b.p, 0, b.cap, s->stream));
    if (words_in > 0)
        B2_CUDA(cudaMemcpyAsync(b.p, words, words_in * 4, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s->stream));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    s->mask_rows[id] = n_rows;
    return B200RAG_OK;
}

int b200rag_mask_set(b200rag_shard* sp, int32_t id, const uint32_t* words, int64_t n_rows) {
    if (sp == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    return mask_store((Shard*)sp, id, words, n_rows, true);
}
int b200rag_mask_set_device(b200rag_shard* sp, int32_t id, const uint32_t* words, int64_t n_rows) {
    if (sp == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    return mask_store((Shard*)sp, id, words, n_rows, false);
}
int b200rag_mask_drop(b200rag_shard* sp, int32_t id) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    auto it = s->masks.find(id);
    if (it == s->masks.end()) return B200RAG_OK;
    use_device(s);
    sync_all(s);
    for (auto& sl : s->slots) sl.staged = false;
    s->staged = false;
    it->second.release();
    s->masks.erase(it);
    s->mask_rows.erase(id);
    return B200RAG_OK;
}

int b200rag_legs_len(const b200rag_query* q, int32_t* nlegs, int32_t* L) {
    if (q == nullptr) { set_error("null query"); return B200RAG_ERR_INVALID; }
    if (nlegs) *nlegs = q->mode == B200RAG_HYBRID ? 2 : 1;
    if (L) *L = q->mode == B200RAG_HYBRID ? 2 * q->top_k : q->top_k;
    return B200RAG_OK;
}

static void activate_slot(Shard* s, const QuerySlot& sl) {
    s->ws.q_bits.p = sl.bits;
    s->ws.q_sp_indptr.p = sl.ind;
    s->ws.q_sp_terms.p = sl.terms;
    s->ws.q_sp_w.p = sl.w;
    s->ws.q_masks.p = sl.masks;
    s->q = sl.q;
    s->h_masks = sl.h_masks;
    s->staged_q_terms = sl.q_terms;
    s->staged = true;
}

int b200rag_stage(b200rag_shard* sp, const b200rag_query* q) { return b200rag_stage_slot(sp, q, 0); }

int b200rag_use_slot(b200rag_shard* sp, int32_t slot) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || slot < 0 || slot >= (int)s->slots.size() || !s->slots[(size_t)slot].staged) {
        set_error("use_slot: no batch staged in this slot");
        return B200RAG_ERR_STATE;
    }
    activate_slot(s, s->slots[(size_t)slot]);
    return B200RAG_OK;
}

static int stage_impl(b200rag_shard* sp, const b200rag_query* q, int32_t slot, bool dev);

int b200rag_stage_slot(b200rag_shard* sp, const b200rag_query* q, int32_t slot) { return stage_impl(sp, q, slot, false); }
int b200rag_stage_device(b200rag_shard* sp, const b200rag_query* q, int32_t slot) { return stage_impl(sp, q, slot, true); }

int b200rag_normalize_bf16_device(b200rag_shard* sp, const float* x, int64_t n, uint16_t* out) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || x == nullptr || out == nullptr || n < 0) { set_error("normalize_device: bad argument"); return B200RAG_ERR_INVALID; }
    if (n == 0) return B200RAG_OK;
    B2_TRY(use_device(s));
    B2_TRY(s->ws.flag.ensure(4, 0, s->stream));
    B2_CUDA(cudaMemsetAsync(s->ws.flag.p, 0, 4, s->stream));
    normalize_bf16_kernel<<<(unsigned)((n + 63) / 64), 64, 0, s->stream>>>(x, n, s->dim, out, s->ws.flag.as<int>());
    B2_CUDA(cudaGetLastError());
    int bad = 0;
    B2_CUDA(cudaMemcpyAsync(&bad, s->ws.flag.p, 4, cudaMemcpyDeviceToHost, s->stream));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    if (bad) { set_error("normalize: non-finite input"); return B200RAG_ERR_INVALID; }
    return B200RAG_OK;
}

static int stage_impl(b200rag_shard* sp, const b200rag_query* q, int32_t slot, bool dev) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || q == nullptr) { set_error("stage: null argument"); return B200RAG_ERR_INVALID; }
    if (slot < 0 || slot >= 4096) { set_error("stage: slot out of range"); return B200RAG_ERR_INVALID; }
    if (q->mode < 0 || q->mode > 2) { set_error("stage: bad mode"); return B200RAG_ERR_INVALID; }
    if (q->batch < 1 || q->batch > 65535) { set_error("stage: batch must be in [1, 65535]"); return B200RAG_ERR_INVALID; }
    if (q->top_k < 1 || q->top_k > B200RAG_MAX_TOPK) { set_error("stage: top_k out of range"); return B200RAG_ERR_INVALID; }
    const int B = q->batch;
    const bool need_dense = q->mode != B200RAG_SPARSE, need_sparse = q->mode != B200RAG_DENSE;
    if (need_dense && q->q_dense_bits == nullptr) { set_error("stage: dense query missing"); return B200RAG_ERR_INVALID; }
    int64_t nt = 0;
    if (need_sparse) {
        if (q->q_sp_indptr == nullptr || q->q_sp_indptr[0] != 0) { set_error("stage: sparse query indptr missing or not starting at 0"); return B200RAG_ERR_INVALID; }
        nt = q->q_sp_indptr[B];
        if (nt > 0 && (q->q_sp_terms == nullptr || q->q_sp_weights == nullptr)) { set_error("stage: sparse query terms missing"); return B200RAG_ERR_INVALID; }
        for (int b = 0; b < B; ++b) {
            if (q->q_sp_indptr[b + 1] < q->q_sp_indptr[b]) { set_error("stage: sparse query indptr not monotone"); return B200RAG_ERR_INVALID; }
            if (dev) continue;       // terms live on the device: validated there, below
            for (int64_t i = q->q_sp_indptr[b]; i < q->q_sp_indptr[b + 1]; ++i) {
                if (q->q_sp_terms[i] >= (uint32_t)s->vocab) { set_error("stage: sparse query index out of vocabulary range"); return B200RAG_ERR_INVALID; }
                if (i > q->q_sp_indptr[b] && q->q_sp_terms[i] <= q->q_sp_terms[i - 1]) { set_error("stage: sparse query indices must be ascending and unique"); return B200RAG_ERR_INVALID; }
            }
        }
    }
    B2_TRY(use_device(s));
    s->h_masks.clear();
    bool any_mask = false;
    if (q->mask_ids != nullptr)
        for (int b = 0; b < B; ++b) any_mask |= q->mask_ids[b] >= 0;
    if (any_mask) {
        s->h_masks.resize((size_t)B, nullptr);
        for (int b = 0; b < B; ++b) {
            const int32_t id = q->mask_ids[b];
            if (id < 0) continue;
            auto it = s->masks.find(id);
            if (it == s->masks.end()) { set_error("stage: unknown mask id"); return B200RAG_ERR_INVALID; }
            if ((size_t)(((s->n_rows + 32767) / 32768) * 32768 / 32) * 4 > it->second.cap) {
                set_error("stage: mask is older than the shard (rows were added since mask_set)");
                return B200RAG_ERR_STATE;
            }
            s->h_masks[(size_t)b] = it->second.as<const uint32_t>();
        }
    }
    // one pinned staging block -> one H2D
    const size_t o_bits = 0;
    const size_t o_ind = al256(o_bits + (need_dense ? (size_t)B * s->dim * 2 : 0));
    const size_t o_terms = al256(o_ind + (size_t)(B + 1) * 8);
    const size_t o_w = al256(o_terms + (size_t)nt * 4);
    const size_t o_masks = al256(o_w + (size_t)nt * 4);
    const size_t total = al256(o_masks + (size_t)B * 8);
    B2_TRY(ensure_pinned(s, total + (size_t)B * q->top_k * 16 + (size_t)(B + 1) * 4 + 1024));
    if ((size_t)slot >= s->slots.size()) s->slots.resize((size_t)slot + 1);
    QuerySlot& sl = s->slots[(size_t)slot];
    sl.staged = false;
    B2_TRY(sl.buf.ensure(total, 0, s->stream));
    // the previous batch's H2D must have drained before the pinned block is rewritten (pipelined mode: a search in
    // flight on the side stream may still read the slot's device block, too)
    B2_TRY(sync_all(s));
    uint8_t* h = (uint8_t*)s->h_pinned;
    uint8_t* d = sl.buf.as<uint8_t>();
    if (need_dense && !dev) memcpy(h + o_bits, q->q_dense_bits, (size_t)B * s->dim * 2);
    if (need_sparse) {
        memcpy(h + o_ind, q->q_sp_indptr, (size_t)(B + 1) * 8);
        if (nt > 0 && !dev) {
            memcpy(h + o_terms, q->q_sp_terms, (size_t)nt * 4);
            memcpy(h + o_w, q->q_sp_weights, (size_t)nt * 4);
        }
    } else {
        memset(h + o_ind, 0, (size_t)(B + 1) * 8);
    }
    if (any_mask) memcpy(h + o_masks, s->h_masks.data(), (size_t)B * 8);
    else memset(h + o_masks, 0, (size_t)B * 8);
    if (!dev) {
        B2_CUDA(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, s->stream));
    } else {
        // host part: indptr + mask pointers; device part: vectors, terms and weights, copied device to device
        B2_CUDA(cudaMemcpyAsync(d + o_ind, h + o_ind, (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, s->stream));
        B2_CUDA(cudaMemcpyAsync(d + o_masks, h + o_masks, (size_t)B * 8, cudaMemcpyHostToDevice, s->stream));
        if (need_dense) B2_CUDA(cudaMemcpyAsync(d + o_bits, q->q_dense_bits, (size_t)B * s->dim * 2, cudaMemcpyDeviceToDevice, s->stream));
        if (need_sparse && nt > 0) {
            B2_CUDA(cudaMemcpyAsync(d + o_terms, q->q_sp_terms, (size_t)nt * 4, cudaMemcpyDeviceToDevice, s->stream));
            B2_CUDA(cudaMemcpyAsync(d + o_w, q->q_sp_weights, (size_t)nt * 4, cudaMemcpyDeviceToDevice, s->stream));
            B2_TRY(s->ws.flag.ensure(4, 0, s->stream));
            B2_CUDA(cudaMemsetAsync(s->ws.flag.p, 0, 4, s->stream));
            validate_sparse_query_kernel<<<B, 128, 0, s->stream>>>((const int64_t*)(d + o_ind), (const uint32_t*)(d + o_terms), B,
                                                                   (uint32_t)s->vocab, s->ws.flag.as<int>());
            B2_CUDA(cudaGetLastError());
            int bad = 0;
            B2_CUDA(cudaMemcpyAsync(&bad, s->ws.flag.p, 4, cudaMemcpyDeviceToHost, s->stream));
            B2_CUDA(cudaStreamSynchronize(s->stream));
            if (bad == 1) { set_error("stage: sparse query index out of vocabulary range"); return B200RAG_ERR_INVALID; }
            if (bad == 2) { set_error("stage: sparse query indices must be ascending and unique"); return B200RAG_ERR_INVALID; }
        }
    }
    sl.bits = d + o_bits;
    sl.ind = d + o_ind;
    sl.terms = d + o_terms;
    sl.w = d + o_w;
    sl.masks = d + o_masks;
    sl.q = *q;
    if (sl.q.rrf_k <= 0) sl.q.rrf_k = 2;
    sl.q.q_dense_bits = nullptr; sl.q.q_sp_indptr = nullptr; sl.q.q_sp_terms = nullptr; sl.q.q_sp_weights = nullptr; sl.q.mask_ids = nullptr;
    sl.h_masks = s->h_masks;
    sl.q_terms = nt;
    sl.staged = true;
    activate_slot(s, sl);
    return B200RAG_OK;
}

int b200rag_legs(b200rag_shard* sp, void* cands_dev, int32_t* ambiguous_dev) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || cands_dev == nullptr) { set_error("legs: null argument"); return B200RAG_ERR_INVALID; }
    if (!s->staged) { set_error("legs: no staged query batch"); return B200RAG_ERR_STATE; }
    B2_TRY(use_device(s));
    s->stats = b200rag_stats{};
    if (s->profile && s->legs_calls > 0) {      // keep the previous call's flags with its event set
        bool* f = s->ev_flags[(s->legs_calls - 1) % kProfileRing];
        f[0] = s->ev_dense; f[1] = s->ev_sparse; f[2] = s->ev_in; f[3] = s->ev_out;
    }
    s->ev = s->ev_ring[s->legs_calls % kProfileRing];
    ++s->legs_calls;
    s->ev_dense = s->ev_sparse = s->ev_in = s->ev_out = false;
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[4], s->stream)); s->ev_in = true; }
    // callers need not pre-zero the counter (pipelined mode: the legs zero it on the stream their tails run on)
    // (the legs zero the ambiguity counter themselves, on the stream their tails run on: run_legs)
    return run_legs(s, (b200rag_cand*)cands_dev, ambiguous_dev);
}

int b200rag_fuse(b200rag_shard* sp, const void* gathered, int32_t n_shards, int32_t has_trailer, int64_t* out_ids,
                 double* out_scores, int32_t* out_counts) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || gathered == nullptr || n_shards < 1 || out_ids == nullptr || out_scores == nullptr || out_counts == nullptr) {
        set_error("fuse: bad argument");
        return B200RAG_ERR_INVALID;
    }
    if (!s->staged) { set_error("fuse: no staged query batch"); return B200RAG_ERR_STATE; }
    B2_TRY(use_device(s));
    const int L = s->q.mode == B200RAG_HYBRID ? 2 * s->q.top_k : s->q.top_k;
    cudaStream_t keep = s->stream;
    if (tail_on_second_stream(s)) s->stream = s->x_stream;       // pipelined mode: the fuse follows the tails
    int rc = launch_fuse(s, s->q.mode, s->q.batch, L, s->q.top_k, s->q.rrf_k, (const b200rag_cand*)gathered, n_shards,
                         has_trailer, out_ids, out_scores, out_counts);
    if (rc == B200RAG_OK && s->profile) {
        if (cudaEventRecord(s->ev[5], s->stream) == cudaSuccess) s->ev_out = true; else rc = cuda_fail(cudaGetLastError(), "cudaEventRecord");
    }
    s->stream = keep;
    if (rc == B200RAG_OK && s->legs_classic) rc = pipeline_handover(s);   // results are read on the second stream
    return rc;
}

int b200rag_search(b200rag_shard* sp, const b200rag_query* q, int64_t* out_ids, double* out_scores,
                   int32_t* out_counts) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || q == nullptr || out_ids == nullptr || out_scores == nullptr || out_counts == nullptr) {
        set_error("search: null argument");
        return B200RAG_ERR_INVALID;
    }
    B2_TRY(b200rag_stage(sp, q));
    // the synchronous host-buffer call gains nothing from the pipelined form: it runs the classic one
    const int saved_pipeline = s->pipeline;
    if (saved_pipeline) { B2_TRY(sync_all(s)); s->pipeline = 0; }
    struct PipeRestore { Shard* s; int v; ~PipeRestore() { s->pipeline = v; } } pipe_restore{s, saved_pipeline};
    cudaStream_t st = s->stream;
    const int B = s->q.batch, K = s->q.top_k;
    const int nlegs = s->q.mode == B200RAG_HYBRID ? 2 : 1;
    const int L = s->q.mode == B200RAG_HYBRID ? 2 * K : K;
    B2_TRY(s->ws.cands.ensure((size_t)nlegs * B * L * sizeof(b200rag_cand), 0, st));
    const size_t o_ids = 0, o_sc = (size_t)B * K * 8, o_cnt = o_sc + (size_t)B * K * 8;
    const size_t out_bytes = o_cnt + (size_t)(B + 1) * 4;
    B2_TRY(s->ws.out.ensure(out_bytes, 0, st));
    uint8_t* d = s->ws.out.as<uint8_t>();
    int32_t* amb = (int32_t*)(d + o_cnt) + B;
    const int saved_slack = s->slack;
    const bool saved_exhaustive = s->exhaustive;
    int retries = 0;
    int launches = 0;
    int rc = B200RAG_OK;
    bool unresolved = false;
    // results land in the tail of the pinned block (the head still feeds the query H2D)
    uint8_t* hres = (uint8_t*)s->h_pinned + (s->h_pinned_cap - al256(out_bytes) - 256);
    hres = (uint8_t*)(((uintptr_t)hres) & ~(uintptr_t)255);
    for (;;) {
        s->stats = b200rag_stats{};
        {
            cudaError_t me = cudaMemsetAsync(amb, 0, 4, st);
            if (me != cudaSuccess) { rc = cuda_fail(me, "cudaMemsetAsync(ambiguous)"); break; }
        }
        rc = run_legs(s, s->ws.cands.as<b200rag_cand>(), amb);
        if (rc != B200RAG_OK) break;
        rc = launch_fuse(s, s->q.mode, B, L, K, s->q.rrf_k, s->ws.cands.as<b200rag_cand>(), 1, 0, (int64_t*)(d + o_ids),
                         (double*)(d + o_sc), (int32_t*)(d + o_cnt));
        if (rc != B200RAG_OK) break;
        cudaError_t e = cudaMemcpyAsync(hres, d, out_bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "result read-back"); break; }
        launches += s->stats.kernel_launches;
        const int32_t ambiguous = ((const int32_t*)(hres + o_cnt))[B];
        if (ambiguous == 0) break;
        const int cur = s->slack > 0 ? s->slack : default_slack(L);
        if (s->exhaustive) { unresolved = true; break; }          // cannot happen: the exhaustive legs never flag
        if (L + cur >= 3 * B200RAG_MAX_TOPK || retries >= 6) {
            // The slack guard never cleared (massive ties / near-duplicate scores around the cut).  Recompute the legs
            // exhaustively -- exact by construction -- or, when that is disabled, refuse to return a guess.
            if (!s->exact_fallback) { unresolved = true; break; }
            s->exhaustive = true;
            ++retries;
            continue;
        }
        s->slack = std::min(cur * 2 + L, 3 * B200RAG_MAX_TOPK - L);  // widen and redo the legs
        ++retries;
    }
    s->slack = saved_slack;
    s->exhaustive = saved_exhaustive;
    if (rc != B200RAG_OK) return rc;
    if (unresolved) {
        set_error("search: the slack guard never cleared (ties or near-duplicate scores around the top-k cut) and the "
                  "exhaustive exact pass is disabled (b200rag_set_exact_fallback / B200RAG_EXACT_FALLBACK)");
        s->stats.retries = retries;
        return B200RAG_ERR_INEXACT;
    }
    memcpy(out_ids, hres + o_ids, (size_t)B * K * 8);
    memcpy(out_scores, hres + o_sc, (size_t)B * K * 8);
    memcpy(out_counts, hres + o_cnt, (size_t)B * 4);
    s->stats.kernel_launches = launches;
    s->stats.retries = retries;
    return B200RAG_OK;
}

}  // extern "C"

// ---- persistence -------------------------------------------------------------------------------------------------
struct ShardFileHeader {
    char magic[8];            // "B200RAG1"
    int32_t version, dim, vocab, reserved;   // version 2 appends the global row ids (i64 [n_rows]) after the weights
    int64_t n_rows, nnz;
};

// Two pinned staging halves: the copy engine fills / drains one half while the host writes / reads the other, so disk and
// PCIe overlap (one event per half; nothing waits for a whole chunk round trip).
struct StagePair {
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    size_t bytes = 0;
    int init(size_t half_bytes) {
        bytes = half_bytes;
        for (int i = 0; i < 2; ++i) {
            if (cudaMallocHost(&buf[i], half_bytes) != cudaSuccess || cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                return B200RAG_ERR_CUDA;
            }
        }
        return B200RAG_OK;
    }
    ~StagePair() {
        for (int i = 0; i < 2; ++i) {
            if (buf[i]) cudaFreeHost(buf[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
    }
};

static int dev_to_file(Shard* s, const void* dev, size_t bytes, FILE* f, StagePair& sp) {
    const size_t nchunks = (bytes + sp.bytes - 1) / sp.bytes;
    auto len = [&](size_t c) { return std::min(sp.bytes, bytes - c * sp.bytes); };
    if (nchunks > 0) {
        B2_CUDA(cudaMemcpyAsync(sp.buf[0], dev, len(0), cudaMemcpyDeviceToHost, s->stream));
        B2_CUDA(cudaEventRecord(sp.ev[0], s->stream));
    }
    for (size_t c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {      // next chunk travels while this one is written
            B2_CUDA(cudaMemcpyAsync(sp.buf[(c + 1) & 1], (const uint8_t*)dev + (c + 1) * sp.bytes, len(c + 1), cudaMemcpyDeviceToHost, s->stream));
            B2_CUDA(cudaEventRecord(sp.ev[(c + 1) & 1], s->stream));
        }
        B2_CUDA(cudaEventSynchronize(sp.ev[c & 1]));
        if (fwrite(sp.buf[c & 1], 1, len(c), f) != len(c)) { set_error("save: short write"); return B200RAG_ERR_INVALID; }
    }
    return B200RAG_OK;
}

static int file_to_dev(Shard* s, void* dev, size_t bytes, FILE* f, StagePair& sp) {
    const size_t nchunks = (bytes + sp.bytes - 1) / sp.bytes;
    auto len = [&](size_t c) { return std::min(sp.bytes, bytes - c * sp.bytes); };
    for (size_t c = 0; c < nchunks; ++c) {
        if (c >= 2) B2_CUDA(cudaEventSynchronize(sp.ev[c & 1]));      // the copy that last used this half has drained
        if (fread(sp.buf[c & 1], 1, len(c), f) != len(c)) { set_error("load: file is truncated"); cudaStreamSynchronize(s->stream); return B200RAG_ERR_INVALID; }
        B2_CUDA(cudaMemcpyAsync((uint8_t*)dev + c * sp.bytes, sp.buf[c & 1], len(c), cudaMemcpyHostToDevice, s->stream));
        B2_CUDA(cudaEventRecord(sp.ev[c & 1], s->stream));
    }
    B2_CUDA(cudaStreamSynchronize(s->stream));
    return B200RAG_OK;
}

extern "C" int b200rag_save(b200rag_shard* sp, const char* path) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || path == nullptr) { set_error("save: null argument"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    B2_TRY(sync_all(s));
    FILE* f = fopen(path, "wb");
    if (f == nullptr) { set_error(std::string("save: cannot open ") + path); return B200RAG_ERR_INVALID; }
    ShardFileHeader h{};
    memcpy(h.magic, "B200RAG1", 8);
    h.version = 2; h.dim = s->dim; h.vocab = s->vocab; h.n_rows = s->n_rows; h.nnz = s->nnz;
    StagePair stage;
    int rc = B200RAG_OK;
    if (stage.init((size_t)32 << 20) != B200RAG_OK) { fclose(f); set_error("save: no pinned staging memory"); return B200RAG_ERR_CUDA; }
    if (fwrite(&h, sizeof(h), 1, f) != 1) { set_error("save: short write"); rc = B200RAG_ERR_INVALID; }
    if (rc == B200RAG_OK && s->n_rows > 0) rc = dev_to_file(s, s->dense.p, (size_t)s->n_rows * s->dim * 2, f, stage);
    if (rc == B200RAG_OK) rc = dev_to_file(s, s->fwd_ptr.p, (size_t)(s->n_rows + 1) * 8, f, stage);
    if (rc == B200RAG_OK && s->nnz > 0) rc = dev_to_file(s, s->fwd_terms.p, (size_t)s->nnz * 4, f, stage);
    if (rc == B200RAG_OK && s->nnz > 0) rc = dev_to_file(s, s->fwd_w.p, (size_t)s->nnz * 4, f, stage);
    if (rc == B200RAG_OK && s->n_rows > 0) rc = dev_to_file(s, s->row_ids.p, (size_t)s->n_rows * 8, f, stage);
    if (fclose(f) != 0 && rc == B200RAG_OK) { set_error("save: close failed"); rc = B200RAG_ERR_INVALID; }
    return rc;
}

extern "C" int b200rag_load(b200rag_shard* sp, const char* path) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || path == nullptr) { set_error("load: null argument"); return B200RAG_ERR_INVALID; }
    if (s->n_rows != 0) { set_error("load: the shard must be empty"); return B200RAG_ERR_STATE; }
    B2_TRY(use_device(s));
    B2_TRY(sync_all(s));
    FILE* f = fopen(path, "rb");
    if (f == nullptr) { set_error(std::string("load: cannot open ") + path); return B200RAG_ERR_INVALID; }
    ShardFileHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "B200RAG1", 8) != 0 || (h.version != 1 && h.version != 2)) {
        fclose(f); set_error("load: not a b200rag shard file"); return B200RAG_ERR_INVALID;
    }
    if (h.dim != s->dim || h.vocab != s->vocab || h.n_rows < 0 || h.nnz < 0 || h.n_rows > 0xFFFFFFF0ll) {
        fclose(f); set_error("load: file was written for another dim/vocab"); return B200RAG_ERR_INVALID;
    }
    StagePair stage;
    if (stage.init((size_t)32 << 20) != B200RAG_OK) { fclose(f); set_error("load: no pinned staging memory"); return B200RAG_ERR_CUDA; }
    cudaStream_t st = s->stream;
    int rc = s->dense.ensure((size_t)std::max<int64_t>(h.n_rows, 1) * s->dim * 2, 0, st);
    if (rc == B200RAG_OK) rc = s->fwd_ptr.ensure((size_t)(h.n_rows + 1) * 8, 0, st);
    if (rc == B200RAG_OK) rc = s->fwd_terms.ensure((size_t)(h.nnz + 1) * 4, 0, st);
    if (rc == B200RAG_OK) rc = s->fwd_w.ensure((size_t)(h.nnz + 1) * 4, 0, st);
    if (rc == B200RAG_OK && h.n_rows > 0) rc = file_to_dev(s, s->dense.p, (size_t)h.n_rows * s->dim * 2, f, stage);
    if (rc == B200RAG_OK) rc = file_to_dev(s, s->fwd_ptr.p, (size_t)(h.n_rows + 1) * 8, f, stage);
    if (rc == B200RAG_OK && h.nnz > 0) rc = file_to_dev(s, s->fwd_terms.p, (size_t)h.nnz * 4, f, stage);
    if (rc == B200RAG_OK && h.nnz > 0) rc = file_to_dev(s, s->fwd_w.p, (size_t)h.nnz * 4, f, stage);
    if (rc == B200RAG_OK) rc = s->row_ids.ensure((size_t)std::max<int64_t>(h.n_rows, 1) * 8, 0, st);
    int64_t last_id = INT64_MIN;
    if (rc == B200RAG_OK && h.n_rows > 0) {
        if (h.version >= 2) rc = file_to_dev(s, s->row_ids.p, (size_t)h.n_rows * 8, f, stage);
        else rc = launch_fill_row_ids(s, s->row_ids.as<int64_t>(), s->cfg.row_base, h.n_rows);
        if (rc == B200RAG_OK) {
            cudaError_t e = cudaMemcpyAsync(&last_id, s->row_ids.as<int64_t>() + (h.n_rows - 1), 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = cuda_fail(e, "load: row ids");
        }
    }
    fclose(f);
    if (rc != B200RAG_OK) {
        // leave a clean empty shard behind: the forward index's first word is what append_rows builds on
        const std::string msg = g_err;
        s->n_rows = 0; s->nnz = 0; s->last_id = INT64_MIN;
        if (s->fwd_ptr.p != nullptr) { cudaMemsetAsync(s->fwd_ptr.p, 0, 8, st); cudaStreamSynchronize(st); cudaGetLastError(); }
        g_err = msg;
        return rc;
    }
    s->n_rows = h.n_rows;
    s->nnz = h.nnz;
    s->last_id = last_id;
    s->built_rows = 0; s->n_blocks = 0; s->inv_nnz = 0; s->w_absmax = 0.f; s->wmax_nnz = 0;
    s->h_blk_base.clear();
    s->q8_rows = 0;
    B2_TRY(q8_sync_rows(s));
    return build_inverted(s);
}

// ---- peer-memory exchange --------------------------------------------------------------------------------------
static void p2p_release(Shard* s) {
    for (int r = 0; r < (int)s->xpeers.size(); ++r)
        if (r != s->x_rank && s->xpeers[(size_t)r] != nullptr) cudaIpcCloseMemHandle(s->xpeers[(size_t)r]);
    s->xpeers.clear();
    if (s->xwin != nullptr) cudaFree(s->xwin);
    s->xwin = nullptr;
    s->ws.xpeers_dev.release();
    s->x_world = 0;
    cudaGetLastError();
}

static size_t p2p_window_bytes(int world, int64_t slot_bytes) {
    return (size_t)2 * world * slot_bytes + (size_t)world * kFlagStrideU64 * 8;
}

extern "C" {

int b200rag_p2p_export(b200rag_shard* sp, int32_t world, int64_t slot_bytes, uint8_t* handle_out) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || handle_out == nullptr || world < 1 || world > 64 || slot_bytes < 16 || slot_bytes % 16 != 0) {
        set_error("p2p_export: bad argument");
        return B200RAG_ERR_INVALID;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == B200RAG_IPC_HANDLE_BYTES, "IPC handle size");
    B2_TRY(use_device(s));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    p2p_release(s);
    const size_t bytes = p2p_window_bytes(world, slot_bytes);
    B2_CUDA(cudaMalloc(&s->xwin, bytes));
    B2_CUDA(cudaMemset(s->xwin, 0, bytes));
    B2_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    B2_CUDA(cudaIpcGetMemHandle(&h, s->xwin));
    memcpy(handle_out, &h, sizeof(h));
    s->x_world = world;
    s->x_slot_bytes = slot_bytes;
    s->x_epoch = 0;
    if (const char* e = getenv("B200RAG_P2P_TIMEOUT_MS")) { const long long ms = atoll(e); if (ms > 0) s->x_timeout_cycles = ms * 2000000ll; }
    return B200RAG_OK;
}

int b200rag_p2p_attach(b200rag_shard* sp, int32_t rank, int32_t world, const uint8_t* handles) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || handles == nullptr || s->xwin == nullptr || world != s->x_world || rank < 0 || rank >= world) {
        set_error("p2p_attach: export a window of the same world size first");
        return B200RAG_ERR_STATE;
    }
    B2_TRY(use_device(s));
    s->x_rank = rank;
    s->xpeers.assign((size_t)world, nullptr);
    for (int r = 0; r < world; ++r) {
        if (r == rank) { s->xpeers[(size_t)r] = s->xwin; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * B200RAG_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { p2p_release(s); return cuda_fail(e, "cudaIpcOpenMemHandle (peer exchange window)"); }
        s->xpeers[(size_t)r] = ptr;
    }
    B2_TRY(s->ws.xpeers_dev.ensure((size_t)world * 8, 0, s->stream));
    B2_CUDA(cudaMemcpyAsync(s->ws.xpeers_dev.p, s->xpeers.data(), (size_t)world * 8, cudaMemcpyHostToDevice, s->stream));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    return B200RAG_OK;
}

int b200rag_p2p_exchange(b200rag_shard* sp, const void* mine, int64_t nbytes) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || mine == nullptr || s->xpeers.empty()) { set_error("p2p_exchange: no attached exchange window"); return B200RAG_ERR_STATE; }
    if (nbytes <= 0 || nbytes % 16 != 0 || nbytes > s->x_slot_bytes) { set_error("p2p_exchange: block does not fit the exchange slot"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    ++s->x_epoch;
    cudaStream_t keep = s->stream;
    if (tail_on_second_stream(s)) s->stream = s->x_stream;
    const int rc = launch_exchange(s, mine, nbytes, s->ws.xpeers_dev.as<void*>(), s->x_world, s->x_rank, s->x_slot_bytes,
                                   (int)(s->x_epoch & 1ull), s->x_epoch);
    s->stream = keep;
    return rc;
}

int b200rag_p2p_set_stream(b200rag_shard* sp, void* stream) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    s->x_stream = (cudaStream_t)stream;
    return B200RAG_OK;
}

int b200rag_p2p_fuse(b200rag_shard* sp, int64_t* out_ids, double* out_scores, int32_t* out_counts) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || out_ids == nullptr || out_scores == nullptr || out_counts == nullptr || s->xpeers.empty()) {
        set_error("p2p_fuse: bad argument or no attached exchange window");
        return B200RAG_ERR_INVALID;
    }
    if (!s->staged || s->x_epoch == 0) { set_error("p2p_fuse: stage + legs + exchange first"); return B200RAG_ERR_STATE; }
    B2_TRY(use_device(s));
    const int L = s->q.mode == B200RAG_HYBRID ? 2 * s->q.top_k : s->q.top_k;
    const uint8_t* win = (const uint8_t*)s->xwin;
    const b200rag_cand* gathered = (const b200rag_cand*)(win + (size_t)(s->x_epoch & 1ull) * s->x_world * s->x_slot_bytes);
    const unsigned long long* flags = (const unsigned long long*)(win + (size_t)2 * s->x_world * s->x_slot_bytes);
    cudaStream_t keep = s->stream;
    if (tail_on_second_stream(s)) s->stream = s->x_stream;
    int rc = launch_fuse(s, s->q.mode, s->q.batch, L, s->q.top_k, s->q.rrf_k, gathered, s->x_world, 1, out_ids, out_scores,
                         out_counts, s->x_slot_bytes / (int64_t)sizeof(b200rag_cand), flags, s->x_epoch);
    if (rc == B200RAG_OK && s->profile) {
        if (cudaEventRecord(s->ev[5], s->stream) == cudaSuccess) s->ev_out = true; else rc = cuda_fail(cudaGetLastError(), "cudaEventRecord");
    }
    s->stream = keep;
    if (rc == B200RAG_OK && s->legs_classic) rc = pipeline_handover(s);   // results are read on the second stream
    return rc;
}

int b200rag_p2p_close(b200rag_shard* sp) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    p2p_release(s);
    return B200RAG_OK;
}

int b200rag_get_stats(const b200rag_shard* sp, b200rag_stats* out) {
    if (sp == nullptr || out == nullptr) { set_error("get_stats: null argument"); return B200RAG_ERR_INVALID; }
    Shard* s = (Shard*)sp;
    if (s->ws.post_count.p != nullptr && s->staged) {
        // not on the timed path: read the postings counter the last sparse scan accumulated
        unsigned long long v = 0;
        if (cudaSetDevice(s->cfg.device) == cudaSuccess && sync_all(s) == B200RAG_OK &&
            cudaMemcpyAsync(&v, s->ws.post_count.p, 8, cudaMemcpyDeviceToHost, s->stream) == cudaSuccess &&
            cudaStreamSynchronize(s->stream) == cudaSuccess)
            s->stats.sparse_postings = (int64_t)v;
    }
    if (s->profile && (s->ev_dense || s->ev_sparse) && sync_all(s) == B200RAG_OK) {
        float ms = 0.f;
        if (s->ev_dense && cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]) == cudaSuccess) s->stats.dense_scan_ms = ms;
        if (s->ev_sparse && cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]) == cudaSuccess) s->stats.sparse_scan_ms = ms;
        if (s->ev_in && s->ev_dense && cudaEventElapsedTime(&ms, s->ev[4], s->ev[0]) == cudaSuccess) s->stats.pre_scan_ms = ms;
        if (s->ev_out && s->ev_dense && cudaEventElapsedTime(&ms, s->ev[1], s->ev[5]) == cudaSuccess) s->stats.tail_ms = ms;
        cudaGetLastError();
    }
    *out = s->stats;
    return B200RAG_OK;
}

int b200rag_get_stats_step(const b200rag_shard* sp, int32_t steps_back, b200rag_stats* out) {
    if (sp == nullptr || out == nullptr || steps_back < 0 || steps_back >= kProfileRing) { set_error("get_stats_step: bad argument"); return B200RAG_ERR_INVALID; }
    Shard* s = (Shard*)sp;
    *out = b200rag_stats{};
    if (!s->profile || s->legs_calls <= steps_back) { set_error("get_stats_step: profiling is off or no such call"); return B200RAG_ERR_STATE; }
    if (cudaSetDevice(s->cfg.device) != cudaSuccess || sync_all(s) != B200RAG_OK) return cuda_fail(cudaGetLastError(), "get_stats_step");
    const int64_t call = s->legs_calls - 1 - steps_back;
    cudaEvent_t* ev = s->ev_ring[call % kProfileRing];
    bool f[4];
    if (steps_back == 0) { f[0] = s->ev_dense; f[1] = s->ev_sparse; f[2] = s->ev_in; f[3] = s->ev_out; }
    else { const bool* g = s->ev_flags[call % kProfileRing]; f[0] = g[0]; f[1] = g[1]; f[2] = g[2]; f[3] = g[3]; }
    float ms = 0.f;
    if (f[0] && cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) out->dense_scan_ms = ms;
    if (f[1] && cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) out->sparse_scan_ms = ms;
    if (f[2] && f[0] && cudaEventElapsedTime(&ms, ev[4], ev[0]) == cudaSuccess) out->pre_scan_ms = ms;
    if (f[3] && f[0] && cudaEventElapsedTime(&ms, ev[1], ev[5]) == cudaSuccess) out->tail_ms = ms;
    cudaGetLastError();
    return B200RAG_OK;
}

int b200rag_set_profiling(b200rag_shard* sp, int32_t on) {
    Shard* s = (Shard*)sp;
    if (s == nullptr) { set_error("null shard"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    if (on && s->ev_ring[0][0] == nullptr)
        for (int r = 0; r < kProfileRing; ++r)
            for (int i = 0; i < 6; ++i) B2_CUDA(cudaEventCreate(&s->ev_ring[r][i]));
    s->profile = on != 0;
    s->ev_dense = s->ev_sparse = s->ev_in = s->ev_out = false;
    return B200RAG_OK;
}

int b200rag_synth_dense(b200rag_shard* sp, uint64_t seed, int64_t row0, int64_t n, uint16_t* out) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || out == nullptr || n < 0) { set_error("synth_dense: bad argument"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    return launch_synth_dense(s->stream, seed, row0, n, s->dim, out);
}

int b200rag_synth_sparse(b200rag_shard* sp, uint64_t seed, int64_t row0, int64_t n, int32_t doc_tokens,
                         const uint64_t* thr, const float* idf, const float* tff, int64_t term_mul, int64_t* counts,
                         const int64_t* indptr, uint32_t* terms, float* weights) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || thr == nullptr || n < 0) { set_error("synth_sparse: bad argument"); return B200RAG_ERR_INVALID; }
    if (counts == nullptr && (indptr == nullptr || terms == nullptr || weights == nullptr || idf == nullptr || tff == nullptr)) {
        set_error("synth_sparse: fill pass needs indptr/terms/weights/idf/tff");
        return B200RAG_ERR_INVALID;
    }
    B2_TRY(use_device(s));
    return launch_synth_sparse(s->stream, seed, row0, n, s->vocab, doc_tokens, thr, idf, tff, term_mul, counts, indptr,
                               terms, weights);
}

int b200rag_exclusive_scan_i64(b200rag_shard* sp, const int64_t* in, int64_t n, int64_t* out) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || in == nullptr || out == nullptr || n < 0) { set_error("scan: bad argument"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    return launch_exclusive_scan_i64(s->stream, in, n, out);
}

int b200rag_synth_collection_mask(b200rag_shard* sp, uint64_t seed, int64_t row0, int64_t n, const uint64_t* thr,
                                  int32_t n_coll, int32_t coll, uint32_t* out_words) {
    Shard* s = (Shard*)sp;
    if (s == nullptr || thr == nullptr || out_words == nullptr || n < 0) { set_error("collection_mask: bad argument"); return B200RAG_ERR_INVALID; }
    B2_TRY(use_device(s));
    return launch_synth_collection_mask(s->stream, seed, row0, n, thr, n_coll, coll, out_words);
}

}  // extern "C"
