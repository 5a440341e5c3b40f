// K3/K4  selection, exact re-score and fusion kernels.
//
//   merge_lists     tree of shared-memory bitonic merges over per-CTA / per-block candidate key lists
//   rescore_dense   canonical fp64 score of the surviving candidates, bit-equal to the oracle
//   rescore_sparse  canonical fp64 sparse score from the forward index
//   finalize_leg    order by (score desc, row asc) (R5), apply score_threshold (R6), slack guard, emit candidates
//   fuse            G-way merge of gathered per-shard legs (R5) + Reciprocal Rank Fusion (R9/R10); replaces
//                   qdrant-client hybrid/fusion.py::reciprocal_rank_fusion as requested at
//                   src/audio_rag/retrieval/qdrant.py:281-298 (FusionQuery(Fusion.RRF), limit=top_k)
#include "common.cuh"
#include "engine.h"

namespace b200rag {

// ------------------------------------------------------------------------------------------------ merge
__global__ void __launch_bounds__(512) merge_lists_kernel(const uint64_t* __restrict__ in, int n_lists, int Lc,
                                                          uint64_t* __restrict__ out, int lists_per_group,
                                                          int n_groups, int npow2) {
    extern __shared__ __align__(16) uint64_t mkeys[];
    const int q = blockIdx.y, g = blockIdx.x;
    const int l0 = g * lists_per_group;
    const int l1 = min(n_lists, l0 + lists_per_group);
    const int cnt = (l1 - l0) * Lc;
    const uint64_t* src = in + ((size_t)q * n_lists + l0) * Lc;
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) mkeys[i] = i < cnt ? src[i] : 0ull;
    cta_bitonic_desc(mkeys, npow2, threadIdx.x, blockDim.x, 0);
    uint64_t* dst = out + ((size_t)q * n_groups + g) * Lc;
    for (int i = threadIdx.x; i < Lc; i += blockDim.x) dst[i] = mkeys[i];
}

int launch_merge_tree(Shard* s, int batch, int n_lists, int Lc, uint64_t* a, uint64_t* b, uint64_t** result) {
    static AttrCache attr;
    if (attr.raise(s->cfg.device, kMergeMaxKeys * 8))
        B2_CUDA(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kMergeMaxKeys * 8));
    // Many small groups per level instead of one big sort: a bitonic network over n keys costs
    // ~log2(n)^2/2 barrier-separated stages, so groups are kept at <= kMergeGroupKeys keys and the tree
    // gets an extra (cheap, parallel) level instead.
    uint64_t* cur = a;
    uint64_t* nxt = b;
    while (n_lists > 1) {
        int lpg = kMergeGroupKeys / Lc;           // ~1024-key groups: more, cheaper levels beat fewer, bigger sorts
        if (lpg < 2) lpg = 2;
        if (lpg > n_lists) lpg = n_lists;
        const int n_groups = (n_lists + lpg - 1) / lpg;
        const int npow2 = next_pow2(lpg * Lc);
        if (npow2 > kMergeMaxKeys) { set_error("merge: candidate list too long"); return B200RAG_ERR_INVALID; }
        int threads = npow2 / 2;
        if (threads < 64) threads = 64;
        if (threads > 512) threads = 512;
        dim3 grid(n_groups, batch);
        merge_lists_kernel<<<grid, threads, (size_t)npow2 * 8, s->stream>>>(cur, n_lists, Lc, nxt, lpg, n_groups, npow2);
        B2_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
        uint64_t* t = cur; cur = nxt; nxt = t;
        n_lists = n_groups;
    }
    *result = cur;
    return B200RAG_OK;
}

// ------------------------------------------------------------------------------------------------ exact dense
__device__ __forceinline__ double bf16lo_d(uint32_t u) { return (double)__uint_as_float(u << 16); }
__device__ __forceinline__ double bf16hi_d(uint32_t u) { return (double)__uint_as_float(u & 0xFFFF0000u); }

// One warp per candidate.  Canonical order (oracle.py::dense_scores): lane l owns elements k with (k % 256) / 8 == l,
// adds its exact fp64 products in ascending k, then the 32 lane partials are combined by a pairwise tree.
__global__ void __launch_bounds__(256) rescore_dense_kernel(const uint16_t* __restrict__ corpus, int dim,
                                                            const uint16_t* __restrict__ q_bits,
                                                            const uint64_t* __restrict__ approx,
                                                            uint64_t* __restrict__ exact, int64_t Lc, int q0) {
    const int q = blockIdx.y;                 // list index; the query it belongs to is q0 + q
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + w;
    if (i >= Lc) return;
    const uint64_t key = approx[(size_t)q * Lc + i];
    if (key == 0) {
        if (lane == 0) exact[(size_t)q * Lc + i] = 0;
        return;
    }
    const uint32_t row = key_row(key);
    const uint4* rp = reinterpret_cast<const uint4*>(corpus + (size_t)row * dim);
    const uint4* qp = reinterpret_cast<const uint4*>(q_bits + (size_t)(q0 + q) * dim);
    const int nch = dim / 256;
    uint4 cv[4], qv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < nch) { cv[c] = rp[c * 32 + lane]; qv[c] = qp[c * 32 + lane]; }
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < nch) {
            acc = __dadd_rn(acc, __dmul_rn(bf16lo_d(cv[c].x), bf16lo_d(qv[c].x)));
            acc = __dadd_rn(acc, __dmul_rn(bf16hi_d(cv[c].x), bf16hi_d(qv[c].x)));
            acc = __dadd_rn(acc, __dmul_rn(bf16lo_d(cv[c].y), bf16lo_d(qv[c].y)));
            acc = __dadd_rn(acc, __dmul_rn(bf16hi_d(cv[c].y), bf16hi_d(qv[c].y)));
            acc = __dadd_rn(acc, __dmul_rn(bf16lo_d(cv[c].z), bf16lo_d(qv[c].z)));
            acc = __dadd_rn(acc, __dmul_rn(bf16hi_d(cv[c].z), bf16hi_d(qv[c].z)));
            acc = __dadd_rn(acc, __dmul_rn(bf16lo_d(cv[c].w), bf16lo_d(qv[c].w)));
            acc = __dadd_rn(acc, __dmul_rn(bf16hi_d(cv[c].w), bf16hi_d(qv[c].w)));
        }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, d));
    if (lane == 0) exact[(size_t)q * Lc + i] = make_key(__double2float_rn(acc) + 0.0f, row);
}

int launch_rescore_dense(Shard* s, int batch, int64_t Lc, const uint64_t* approx, uint64_t* exact, int q0) {
    dim3 grid((unsigned)((Lc + 7) / 8), batch);
    rescore_dense_kernel<<<grid, 256, 0, s->stream>>>(s->dense.as<uint16_t>(), s->dim,
                                                      s->ws.q_bits.as<uint16_t>(), approx, exact, Lc, q0);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

// Exact sparse score of one document against one chunk of query terms (<= kMaxQueryTermsChunk, ascending, in shared
// memory), by one warp.  The DOCUMENT's terms and weights are streamed with coalesced loads, 256 per round and all of a
// round's loads in flight at once (~200 terms per 256-token chunk: one round), and each term is looked up in the query
// chunk by binary search in shared memory -- one global round trip instead of a chain of dependent loads per query term.
// Products land in prod[j] (j = position of the term in the query), lane 0 then adds them in ascending j = ascending
// term index: the canonical order of SURVEY R3, bit-equal to the oracle.
__device__ __forceinline__ void sparse_exact_chunk(int64_t ds, int64_t de, const uint32_t* __restrict__ fwd_terms,
                                                   const float* __restrict__ fwd_w, const uint32_t* qt, const float* qw,
                                                   int cn, double* prod, uint8_t* present, int lane, double& acc,
                                                   int& touched) {
    for (int j = lane; j < cn; j += 32) present[j] = 0;
    __syncwarp();
    const uint32_t tlo = qt[0], thi = qt[cn - 1];
    for (int64_t i0 = ds; i0 < de; i0 += 256) {      // 256 document terms per round: 8 term + 8 weight loads in flight per lane
        uint32_t t8[8];
        float w8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t i = i0 + u * 32 + lane;
            t8[u] = i < de ? fwd_terms[i] : 0xFFFFFFFFu;
            w8[u] = i < de ? fwd_w[i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t t = t8[u];
            if (t < tlo || t > thi) continue;
            int lo = 0, hi = cn;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (qt[mid] < t) lo = mid + 1; else hi = mid;
            }
            if (lo < cn && qt[lo] == t) {
                present[lo] = 1;
                prod[lo] = __dmul_rn((double)qw[lo], (double)w8[u]);
            }
        }
    }
    __syncwarp();
    if (lane == 0)
        for (int j = 0; j < cn; ++j)
            if (present[j]) { acc = __dadd_rn(acc, prod[j]); touched = 1; }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------ exact sparse
__global__ void __launch_bounds__(256) rescore_sparse_kernel(const int64_t* __restrict__ fwd_ptr,
                                                             const uint32_t* __restrict__ fwd_terms,
                                                             const float* __restrict__ fwd_w,
                                                             const int64_t* __restrict__ q_indptr,
                                                             const uint32_t* __restrict__ q_terms,
                                                             const float* __restrict__ q_w,
                                                             const uint64_t* __restrict__ approx,
                                                             uint64_t* __restrict__ exact, int64_t Lc, int q0,
                                                             int drop_untouched) {
    __shared__ uint32_t qt[kMaxQueryTermsChunk];
    __shared__ float qw[kMaxQueryTermsChunk];
    __shared__ double prod[8][kMaxQueryTermsChunk];
    __shared__ uint8_t present[8][kMaxQueryTermsChunk];
    const int q = blockIdx.y;                 // list index; the query it belongs to is q0 + q
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + w;
    const uint64_t key = i < Lc ? approx[(size_t)q * Lc + i] : 0ull;
    const bool active = key != 0;
    const uint32_t row = key_row(key);
    int64_t ds = 0, de = 0;
    if (active) { ds = fwd_ptr[row]; de = fwd_ptr[row + 1]; }
    const int64_t qs = q_indptr[q0 + q], qe = q_indptr[q0 + q + 1];
    double acc = 0.0;
    int touched = 0;
    for (int64_t c0 = qs; c0 < qe; c0 += kMaxQueryTermsChunk) {
        const int cn = (int)min((int64_t)kMaxQueryTermsChunk, qe - c0);
        __syncthreads();
        for (int j = threadIdx.x; j < cn; j += blockDim.x) { qt[j] = q_terms[c0 + j]; qw[j] = q_w[c0 + j]; }
        __syncthreads();
        if (active) sparse_exact_chunk(ds, de, fwd_terms, fwd_w, qt, qw, cn, prod[w], present[w], lane, acc, touched);
    }
    if (i < Lc && lane == 0)
        exact[(size_t)q * Lc + i] = (active && (touched || !drop_untouched)) ? make_key(__double2float_rn(acc) + 0.0f, row) : 0ull;
}

int launch_rescore_sparse(Shard* s, int batch, int64_t Lc, const uint64_t* approx, uint64_t* exact, int q0,
                          bool drop_untouched) {
    dim3 grid((unsigned)((Lc + 7) / 8), batch);
    rescore_sparse_kernel<<<grid, 256, 0, s->stream>>>(s->fwd_ptr.as<int64_t>(), s->fwd_terms.as<uint32_t>(),
                                                       s->fwd_w.as<float>(), s->ws.q_sp_indptr.as<int64_t>(),
                                                       s->ws.q_sp_terms.as<uint32_t>(), s->ws.q_sp_w.as<float>(),
                                                       approx, exact, Lc, q0, drop_untouched ? 1 : 0);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

// ------------------------------------------------------------------------------------------------ finalize leg
__global__ void __launch_bounds__(256) finalize_leg_kernel(const uint64_t* __restrict__ approx,
                                                           const uint64_t* __restrict__ exact, int Lc, int L,
                                                           int npow2, float eps_abs, float eps_rel,
                                                           const float* __restrict__ eps_abs_q, int has_thr,
                                                           float thr, const int64_t* __restrict__ row_ids,
                                                           b200rag_cand* __restrict__ out,
                                                           int32_t* __restrict__ ambiguous) {
    extern __shared__ __align__(16) uint64_t fkeys[];
    __shared__ int nvalid_s;
    const int q = blockIdx.x;
    if (threadIdx.x == 0) nvalid_s = 0;
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        uint64_t k = i < Lc ? exact[(size_t)q * Lc + i] : 0ull;
        if (k != 0 && has_thr && key_score(k) < thr) k = 0;
        fkeys[i] = k;
    }
    cta_bitonic_desc(fkeys, npow2, threadIdx.x, blockDim.x, 0);
    int local = 0;
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) local += fkeys[i] != 0 ? 1 : 0;
    if (local) atomicAdd(&nvalid_s, local);
    __syncthreads();
    const int nvalid = nvalid_s;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const uint64_t k = fkeys[i];
        b200rag_cand c;
        c.id = k != 0 ? row_ids[key_row(k)] : -1;
        c.score = k != 0 ? key_score(k) : 0.f;
        c.valid = k != 0 ? 1u : 0u;
        out[(size_t)q * L + i] = c;
    }
    if (threadIdx.x == 0 && ambiguous != nullptr) {
        const uint64_t last = approx[(size_t)q * Lc + Lc - 1];
        if (last != 0) {  // the approximate list was full: rows outside it exist, bounded by `a`
            const float a = key_score(last);
            bool have_bound = false;
            float bound = 0.f;
            if (nvalid >= L) { bound = key_score(fkeys[L - 1]); have_bound = true; }
            else if (has_thr) { bound = thr; have_bound = true; }
            if (!have_bound) {
                atomicAdd(ambiguous, 1);  // fewer than L survivors although candidates were cut: widen
            } else {
                const float eps = eps_abs + (eps_abs_q != nullptr ? eps_abs_q[q] : 0.f) + eps_rel * fmaxf(fabsf(a), fabsf(bound));
                if (a + eps >= bound) atomicAdd(ambiguous, 1);
            }
        }
    }
}

int launch_finalize_leg(Shard* s, int batch, int Lc, int L, const uint64_t* approx, const uint64_t* exact,
                        float eps_abs, float eps_rel, const float* eps_abs_q, int has_thr, float thr,
                        b200rag_cand* out, int32_t* ambiguous) {
    const int npow2 = next_pow2(Lc);
    finalize_leg_kernel<<<batch, 256, (size_t)npow2 * 8, s->stream>>>(approx, exact, Lc, L, npow2, eps_abs, eps_rel,
                                                                      eps_abs_q, has_thr, thr, s->row_ids.as<int64_t>(), out, ambiguous);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

// ------------------------------------------------------------------------------------------------ fused leg tail
// merge_lists (all levels) + rescore + finalize_leg in ONE launch per leg (n_lists * Lc <= kTailMaxKeys).  After the scan every further launch is pure latency (4-5 launches of 5-15 us each
// plus the gaps between them), which is what caps strong scaling on small shards.
//   1. every thread takes a strided slice of the n_lists * Lc keys and keeps its maximum; each warp sorts its 32 maxima
//      in registers and publishes its k-th largest, k = ceil(Lc / #warps); the minimum over the warps has >= Lc keys at
//      or above it, so only those survivors (typically a few Lc) are compacted and bitonic-sorted
//   2. the best Lc are re-scored exactly, one warp per candidate (same canonical order as the standalone kernels)
//   3. exact keys are sorted, thresholded, guarded and emitted exactly like finalize_leg_kernel
// (kTailMaxKeys, engine.h: the lists stay in global memory -- two strided passes; only survivors go to smem)
constexpr int kTailSurvivorCap = 4096;

__device__ __forceinline__ uint64_t rescore_dense_warp(const uint16_t* __restrict__ corpus, int dim,
                                                       const uint16_t* __restrict__ q_bits, int q, uint32_t row, int lane) {
    const uint4* rp = reinterpret_cast<const uint4*>(corpus + (size_t)row * dim);
    const uint4* qp = reinterpret_cast<const uint4*>(q_bits + (size_t)q * dim);
    const int nch = dim / 256;
    uint4 cv[4], qv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < nch) { cv[c] = rp[c * 32 + lane]; qv[c] = qp[c * 32 + lane]; }
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < nch) {
            acc = __dadd_rn(acc, __dmul_rn(bf16lo_d(cv[c].x), bf16lo_d(qv[c].x)));
            acc = __dadd_rn(acc, __dmul_rn(bf16hi_d(cv[c].x), bf16hi_d(qv[c].x)));
            acc = __dadd_rn(acc, __dmul_rn(bf16lo_d(cv[c].y), bf16lo_d(qv[c].y)));
            acc = __dadd_rn(acc, __dmul_rn(bf16hi_d(cv[c].y), bf16hi_d(qv[c].y)));
            acc = __dadd_rn(acc, __dmul_rn(bf16lo_d(cv[c].z), bf16lo_d(qv[c].z)));
            acc = __dadd_rn(acc, __dmul_rn(bf16hi_d(cv[c].z), bf16hi_d(qv[c].z)));
            acc = __dadd_rn(acc, __dmul_rn(bf16lo_d(cv[c].w), bf16lo_d(qv[c].w)));
            acc = __dadd_rn(acc, __dmul_rn(bf16hi_d(cv[c].w), bf16hi_d(qv[c].w)));
        }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, d));
    return make_key(__double2float_rn(acc) + 0.0f, row);
}

struct TailParams {
    const uint64_t* lists;   // [batch][n_lists][Lc]
    const uint64_t* gkey;    // [batch] or null: the scan's grid-wide threshold key (>= Lc listed keys are at or above it)
    int n_lists, Lc, L;
    // dense re-score
    const uint16_t* corpus;
    int dim;
    const uint16_t* q_bits;
    // sparse re-score
    const int64_t* fwd_ptr;
    const uint32_t* fwd_terms;
    const float* fwd_w;
    const int64_t* q_indptr;
    const uint32_t* q_terms;
    const float* q_w;
    // finalize
    float eps_abs, eps_rel;
    const float* eps_abs_q;
    int has_thr;
    float thr;
    const int64_t* row_ids;  // [n_rows] global id of a local row
    b200rag_cand* out;       // [batch][L]
    int32_t* ambiguous;
};

template <bool SPARSE, int NT>
__global__ void __launch_bounds__(NT) leg_tail_kernel(const TailParams p) {
    extern __shared__ __align__(16) uint64_t tkeys[];     // [kTailSurvivorCap] survivors, later [npow2(Lc)] exact keys
    constexpr int NW = NT / 32;
    __shared__ uint64_t wk_s[NW];
    __shared__ int cnt_s, nvalid_s;
    __shared__ uint64_t last_s;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = p.n_lists * p.Lc;
    const uint64_t* src = p.lists + (size_t)q * total;
    if (tid == 0) { cnt_s = 0; nvalid_s = 0; }

    // ---- 1. threshold, survivors, sort
    // Threshold from the thread maxima: every thread keeps the maximum of its strided slice, every warp sorts its 32
    // maxima and publishes its k-th largest, k = ceil(Lc / #warps); the minimum over the warps has >= Lc keys at or above
    // it.  The scans' own grid-wide threshold (p.gkey: the Lc-th best key of SOME list) is a second valid bound -- a much
    // weaker one (at 10M rows it let 12 000 of 31 700 keys through where the maxima let ~100 through), but the only one
    // when Lc exceeds 32 keys per warp -- so the tighter of the two is used.
    // (The lists are SORTED: with a plain stride of NT a thread would meet the same rank of every list whenever Lc divides
    //  NT -- Lc = 256: thread t only ever sees rank t mod 256 -- the maxima would be stratified by rank, the threshold as
    //  weak as the lists' tails, and nearly every key would survive: 250 us of radix select instead of 60 us.  Rotating
    //  the lane assignment by 37 slots per round gives every thread a spread of ranks and keeps the loads coalesced.)
    // (More than 32 keys per warp to guarantee -- Lc = 768 on 16 warps: the rounds are dealt to C classes, every class
    //  guarantees ceil(Lc / C) keys by the same argument, and the weakest class threshold covers Lc.)
    {
        const int kk = (p.Lc + NW - 1) / NW;
        const int C = (kk + 31) / 32;                      // 1 unless Lc > 32 * #warps
        const int kc = (kk + C - 1) / C;                   // keys per warp and class, <= 32
        uint64_t wmin = ~0ull;
        for (int cl = 0; cl < C; ++cl) {
            uint64_t tmax = 0;
            int j = cl;
            for (int base = cl * NT; base < total; base += 8 * C * NT, j += 8 * C) {   // eight independent loads in flight
                uint64_t k8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = base + u * C * NT + ((tid + 37 * (j + u * C)) & (NT - 1));
                    k8[u] = i < total ? src[i] : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) tmax = k8[u] > tmax ? k8[u] : tmax;
            }
            uint64_t v = tmax;
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
                for (int jj = k >> 1; jj > 0; jj >>= 1) {
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, v, jj);
                    const bool keep_max = (((lane & jj) == 0) == ((lane & k) == 0));
                    v = keep_max ? (o > v ? o : v) : (o < v ? o : v);
                }
            const uint64_t kth = __shfl_sync(0xffffffffu, v, kc - 1);
            wmin = kth < wmin ? kth : wmin;
        }
        if (lane == 0) wk_s[warp] = wmin;
    }
    __syncthreads();
    uint64_t tau = wk_s[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) tau = wk_s[w] < tau ? wk_s[w] : tau;
    {
        const uint64_t gk = p.gkey != nullptr ? p.gkey[q] : 0ull;
        tau = gk > tau ? gk : tau;
    }
    if (tau == 0) tau = 1;                 // fewer than Lc real keys: keep every non-empty slot
    // survivors: eight keys per thread and round in flight, one shared atomic per warp and key group (warp-uniform trips)
    for (int i0 = warp * 256; i0 < total; i0 += NT * 8) {
        uint64_t k4[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 32 + lane;
            k4[u] = i < total ? src[i] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool take = k4[u] >= tau;
            const unsigned m = __ballot_sync(0xffffffffu, take);
            if (m != 0) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&cnt_s, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                const int pos = base + __popc(m & ((1u << lane) - 1u));
                if (take && pos < kTailSurvivorCap) tkeys[pos] = k4[u];
            }
        }
    }
    __syncthreads();
    int M = cnt_s;
    if (M > kTailSurvivorCap) {
        // Too many survivors (massive ties, or no usable threshold over very many lists): radix-select the exact Lc-th
        // largest KEY -- six histogram passes over the 64 key bits (11 + 11 + 10 score bits, 11 + 11 + 10 row bits)
        // instead of one counting pass per bit -- then keep exactly the keys at or above it (keys are unique).
        int* hist = reinterpret_cast<int*>(tkeys);         // 2048 bins; the overflowed survivor area is free
        __shared__ int need_s;
        __shared__ unsigned long long pref_s;
        unsigned long long prefix = 0;
        int need = p.Lc, hi = 64;
        __syncthreads();
#pragma unroll 1
        for (int d = 0; d < 6; ++d) {
            const int w = (d % 3 == 2) ? 10 : 11;
            const int shift = hi - w;
            for (int i = tid; i < (1 << w); i += NT) hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < total; i += NT) {
                const unsigned long long k = src[i];
                if (k != 0 && (hi == 64 || (k >> hi) == prefix)) atomicAdd(&hist[(int)((k >> shift) & ((1ull << w) - 1ull))], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int acc = 0, b = (1 << w) - 1;
                for (; b > 0; --b) {
                    if (acc + hist[b] >= need) break;
                    acc += hist[b];
                }
                need_s = need - acc;
                pref_s = (prefix << w) | (unsigned long long)b;
            }
            __syncthreads();
            need = need_s;
            prefix = pref_s;
            hi = shift;
            __syncthreads();
        }
        const uint64_t K = prefix;
        if (tid == 0) cnt_s = 0;
        __syncthreads();
        for (int i = tid; i < total; i += NT) {
            const uint64_t k = src[i];
            if (k >= K && k != 0) {
                const int pos = atomicAdd(&cnt_s, 1);
                if (pos < kTailSurvivorCap) tkeys[pos] = k;
            }
        }
        __syncthreads();
        M = cnt_s < kTailSurvivorCap ? cnt_s : kTailSurvivorCap;
    }
    int npow2 = next_pow2(M > p.Lc ? M : p.Lc);
    for (int i = M + tid; i < npow2; i += NT) tkeys[i] = 0;
    cta_bitonic_desc(tkeys, npow2, tid, NT, 0);
    if (tid == 0) last_s = tkeys[p.Lc - 1];      // weakest retained approximate key (0: nothing was cut)
    __syncthreads();

    // ---- 2. exact re-score of the best Lc, in place
    if constexpr (!SPARSE) {
        for (int i = warp; i < p.Lc; i += NW) {
            const uint64_t key = tkeys[i];
            uint64_t ex = 0;
            if (key != 0) ex = rescore_dense_warp(p.corpus, p.dim, p.q_bits, q, key_row(key), lane);
            __syncwarp();
            if (lane == 0) tkeys[i] = ex;
        }
    } else {
        // dynamic shared memory behind the survivor area:
        //   prod [NW][chunk] f64 | qt [chunk] u32 | qw [chunk] f32 | present [NW][chunk] u8
        double* prod_base = reinterpret_cast<double*>(tkeys + kTailSurvivorCap);
        uint32_t* qt = reinterpret_cast<uint32_t*>(prod_base + (size_t)NW * kMaxQueryTermsChunk);
        float* qw = reinterpret_cast<float*>(qt + kMaxQueryTermsChunk);
        uint8_t* present_base = reinterpret_cast<uint8_t*>(qw + kMaxQueryTermsChunk);
        double* prod_w = prod_base + (size_t)warp * kMaxQueryTermsChunk;
        uint8_t* present_w = present_base + (size_t)warp * kMaxQueryTermsChunk;
        const int64_t qs = p.q_indptr[q], qe = p.q_indptr[q + 1];
        for (int i0 = 0; i0 < p.Lc; i0 += NW) {
            const int i = i0 + warp;
            const uint64_t key = i < p.Lc ? tkeys[i] : 0ull;
            const bool active = key != 0;
            const uint32_t row = key_row(key);
            int64_t ds = 0, de = 0;
            if (active) { ds = p.fwd_ptr[row]; de = p.fwd_ptr[row + 1]; }
            double acc = 0.0;
            for (int64_t c0 = qs; c0 < qe; c0 += kMaxQueryTermsChunk) {
                const int cn = (int)min((int64_t)kMaxQueryTermsChunk, qe - c0);
                __syncthreads();
                for (int j = tid; j < cn; j += NT) { qt[j] = p.q_terms[c0 + j]; qw[j] = p.q_w[c0 + j]; }
                __syncthreads();
                if (active) {
                    int touched = 0;
                    sparse_exact_chunk(ds, de, p.fwd_terms, p.fwd_w, qt, qw, cn, prod_w, present_w, lane, acc, touched);
                }
            }
            if (i < p.Lc && lane == 0) tkeys[i] = active ? make_key(__double2float_rn(acc) + 0.0f, row) : 0ull;
        }
    }
    __syncthreads();

    // ---- 3. finalize: threshold, order, guard, emit
    const int fpow2 = next_pow2(p.Lc);
    for (int i = tid; i < fpow2; i += NT) {
        uint64_t k = i < p.Lc ? tkeys[i] : 0ull;
        if (k != 0 && p.has_thr && key_score(k) < p.thr) k = 0;
        tkeys[i] = k;
    }
    cta_bitonic_desc(tkeys, fpow2, tid, NT, 0);
    int local = 0;
    for (int i = tid; i < fpow2; i += NT) local += tkeys[i] != 0 ? 1 : 0;
    if (local) atomicAdd(&nvalid_s, local);
    __syncthreads();
    const int nvalid = nvalid_s;
    for (int i = tid; i < p.L; i += NT) {
        const uint64_t k = tkeys[i];
        b200rag_cand c;
        c.id = k != 0 ? p.row_ids[key_row(k)] : -1;
        c.score = k != 0 ? key_score(k) : 0.f;
        c.valid = k != 0 ? 1u : 0u;
        p.out[(size_t)q * p.L + i] = c;
    }
    if (tid == 0 && p.ambiguous != nullptr) {
        const uint64_t last = last_s;
        if (last != 0) {  // the approximate list was full: rows outside it exist, bounded by `a`
            const float a = key_score(last);
            bool have_bound = false;
            float bound = 0.f;
            if (nvalid >= p.L) { bound = key_score(tkeys[p.L - 1]); have_bound = true; }
            else if (p.has_thr) { bound = p.thr; have_bound = true; }
            if (!have_bound) {
                atomicAdd(p.ambiguous, 1);
            } else {
                const float eps = p.eps_abs + (p.eps_abs_q != nullptr ? p.eps_abs_q[q] : 0.f) +
                                  p.eps_rel * fmaxf(fabsf(a), fabsf(bound));
                if (a + eps >= bound) atomicAdd(p.ambiguous, 1);
            }
        }
    }
}

bool leg_tail_fits(int n_lists, int Lc) { return (int64_t)n_lists * Lc <= kTailMaxKeys; }

int launch_leg_tail(Shard* s, bool sparse, int batch, int n_lists, int Lc, int L, const uint64_t* lists, float eps_abs,
                    float eps_rel, const float* eps_abs_q, int has_thr, float thr, b200rag_cand* out,
                    int32_t* ambiguous, const uint64_t* gkey) {
    TailParams p{};
    p.lists = lists; p.gkey = gkey; p.n_lists = n_lists; p.Lc = Lc; p.L = L;
    p.corpus = s->dense.as<uint16_t>(); p.dim = s->dim; p.q_bits = s->ws.q_bits.as<uint16_t>();
    p.fwd_ptr = s->fwd_ptr.as<int64_t>(); p.fwd_terms = s->fwd_terms.as<uint32_t>(); p.fwd_w = s->fwd_w.as<float>();
    p.q_indptr = s->ws.q_sp_indptr.as<int64_t>(); p.q_terms = s->ws.q_sp_terms.as<uint32_t>(); p.q_w = s->ws.q_sp_w.as<float>();
    p.eps_abs = eps_abs; p.eps_rel = eps_rel; p.eps_abs_q = eps_abs_q; p.has_thr = has_thr; p.thr = thr;
    p.row_ids = s->row_ids.as<int64_t>(); p.out = out; p.ambiguous = ambiguous;
    const size_t smem = (size_t)kTailSurvivorCap * 8;
    auto sparse_smem = [&](int nw) { return smem + (size_t)nw * kMaxQueryTermsChunk * 9 + (size_t)kMaxQueryTermsChunk * 8; };
    static AttrCache attr;
    if (attr.raise(s->cfg.device, sparse_smem(32))) {
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sparse_smem(8)));
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sparse_smem(32)));
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sparse_smem(16)));
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<true, 512>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // the sparse leg's tail runs beside the dense scan: same carve-out, or it would wait for the scan's SMs to drain
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<true, 256>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<true, 1024>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<false, 512>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        B2_CUDA(cudaFuncSetAttribute(leg_tail_kernel<false, 1024>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    // More than 64 candidates per leg (top-k beyond ~20): more warps, so that the per-warp k-th-largest threshold works
    // (k = ceil(Lc / #warps) <= 32) and more candidates are re-scored at a time -- 32 warps when the tail runs alone,
    // 16 when it must CO-RESIDE with a dense scan (s->tail_beside_scan, set by run_legs: every tail of the pipelined form,
    // and the sparse leg's tail whenever that leg overlaps the dense scan): a 1024-thread CTA does not fit the register
    // file beside a scan CTA; it would wait for the whole scan to drain and run after it (measured at 12.5M rows,
    // top-100: 4.9 ms per step instead of 4.1).
    const bool beside_scan = s->tail_beside_scan;
    if (sparse && Lc > 64 && !beside_scan) leg_tail_kernel<true, 1024><<<batch, 1024, sparse_smem(32), s->stream>>>(p);
    else if (sparse && Lc > 64) leg_tail_kernel<true, 512><<<batch, 512, sparse_smem(16), s->stream>>>(p);
    else if (sparse) leg_tail_kernel<true, 256><<<batch, 256, sparse_smem(8), s->stream>>>(p);
    else if (Lc > 64 && !beside_scan) leg_tail_kernel<false, 1024><<<batch, 1024, smem, s->stream>>>(p);
    else leg_tail_kernel<false, 512><<<batch, 512, smem, s->stream>>>(p);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

// ------------------------------------------------------------------------------------------------ fuse
__device__ __forceinline__ bool cand_better(const b200rag_cand& f, const b200rag_cand& e) {
    const uint32_t fo = ord_f32(f.score), eo = ord_f32(e.score);
    return fo > eo || (fo == eo && f.id < e.id);
}

// Gathered candidates may have been written by PEER GPUs (exchange windows) moments ago: read them with ld.global.cg
// (L2, which is coherent with NVLink writes) -- never through the non-coherent / L1 path a const __restrict__ pointer
// would otherwise be allowed to take.
__device__ __forceinline__ b200rag_cand ld_cand_cg(const b200rag_cand* p) {
    static_assert(sizeof(b200rag_cand) == 16, "candidate is one 16-byte word pair");
    const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(p));
    b200rag_cand c;
    memcpy(&c, &v, 16);
    return c;
}

// 1 if some shard's block of `stage` ([n_shards][L]) is not in leg order (valid first, then cand_better); per thread
__device__ __forceinline__ int block_unsorted(const b200rag_cand* stage, int M, int L) {
    int bad = 0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        if (i % L == 0) continue;
        const b200rag_cand e = stage[i], pv = stage[i - 1];
        if (e.valid && !(pv.valid && cand_better(pv, e))) bad = 1;
    }
    return bad;
}

// smem layout: stage[M] cands | leg_id[2][L] i64 | leg_score[2L] f32 | fused[2L] f64 | fid[2L] i64 | ford[2L] int | rnk[2L] int
__global__ void __launch_bounds__(1024) fuse_kernel(const b200rag_cand* __restrict__ gathered, int n_shards,
                                                   int64_t shard_stride, int has_trailer, int nlegs, int batch,
                                                   int L, int top_k, int rrf_k,
                                                   int64_t* __restrict__ out_ids, double* __restrict__ out_scores,
                                                   int32_t* __restrict__ out_counts,
                                                   const unsigned long long* wait_flags, unsigned long long wait_epoch,
                                                   long long timeout_cycles) {
    extern __shared__ __align__(16) uint8_t fsm[];
    const int M = n_shards * L;
    b200rag_cand* stage = reinterpret_cast<b200rag_cand*>(fsm);
    int64_t* leg_id = reinterpret_cast<int64_t*>(stage + next_pow2(M));
    float* leg_score = reinterpret_cast<float*>(leg_id + 2 * L);
    double* fused = reinterpret_cast<double*>(leg_score + 2 * L);
    int64_t* fid = reinterpret_cast<int64_t*>(fused + 2 * L);
    int* ford = reinterpret_cast<int*>(fid + 2 * L);
    int* rnk = ford + 2 * L;
    __shared__ int leg_n[2];
    __shared__ int total_s;
    const int q = blockIdx.x;
    if (threadIdx.x < 2) leg_n[threadIdx.x] = 0;
    if (threadIdx.x == 0) total_s = 0;
    if (wait_flags != nullptr) {
        // peer-memory exchange: every shard's block is complete once its flag carries this search's epoch
        // out_counts[batch + 1] is a STICKY latch (zeroed once by the caller, set by any block that ever gave up on a
        // peer): a block that finds it set does not wait at all, so one timeout empties every query of this and of all
        // later searches instead of leaving a mix of fused and emptied queries (the ranks' epochs have diverged)
        __shared__ int timed_out;
        if (threadIdx.x == 0) timed_out = *reinterpret_cast<volatile int32_t*>(&out_counts[batch + 1]) != 0 ? 1 : 0;
        __syncthreads();
        if (timed_out) {
            if (threadIdx.x == 0) { out_counts[q] = 0; if (q == 0) out_counts[batch] = -1; }
            return;
        }
        if (threadIdx.x < n_shards) {
            const unsigned long long* f = wait_flags + (size_t)threadIdx.x * kFlagStrideU64;
            const long long t0 = clock64();
            for (;;) {
                unsigned long long v;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
                if (v >= wait_epoch) break;
                if (clock64() - t0 > timeout_cycles) { timed_out = 1; break; }   // a peer is gone
            }
        }
        __syncthreads();
        if (timed_out) {
            if (threadIdx.x == 0) {
                atomicExch(&out_counts[batch + 1], 1);
                out_counts[q] = 0;
                if (q == 0) out_counts[batch] = -1;
            }
            return;
        }
    }
    if (has_trailer && q == 0 && threadIdx.x == 0) {
        int amb = 0;
        for (int sh = 0; sh < n_shards; ++sh)
            amb += (int)(ld_cand_cg(gathered + (size_t)sh * shard_stride + (size_t)nlegs * batch * L).id & 0xFFFFFFFFll);
        out_counts[batch] = amb;
    }

    // G-way merge of a leg = sort of the gathered candidates under R5 (valid first, score desc, id asc).  Small sets
    // (one or two shards, top-10) are ranked by counting; larger ones (8 shards x top-100 = 1600 per leg, where
    // counting costs M^2 = 2.6 M compares and a millisecond) by a bitonic network over the 16-byte candidates.
    const int Mp = next_pow2(M);
    for (int leg = 0; leg < nlegs; ++leg) {
        __syncthreads();
        for (int i = threadIdx.x; i < Mp; i += blockDim.x) {
            if (i < M) {
                const int sh = i / L, j = i - sh * L;
                stage[i] = ld_cand_cg(gathered + (size_t)sh * shard_stride + (((size_t)leg * batch + q) * L + j));
            } else {
                b200rag_cand e; e.id = -1; e.score = 0.f; e.valid = 0u;
                stage[i] = e;
            }
        }
        __syncthreads();
        if (M <= 128) {
            int local = 0;
            for (int i = threadIdx.x; i < M; i += blockDim.x) {
                const b200rag_cand e = stage[i];
                if (!e.valid) continue;
                ++local;
                int rank = 0;
                for (int f = 0; f < M; ++f) {
                    const b200rag_cand o = stage[f];
                    if (o.valid && cand_better(o, e)) ++rank;
                }
                if (rank < L) { leg_id[leg * L + rank] = e.id; leg_score[leg * L + rank] = e.score; }
            }
            if (local) atomicAdd(&leg_n[leg], local);
        } else if (__syncthreads_or(block_unsorted(stage, M, L)) == 0) {
            // every shard's block arrives sorted under R5 (the leg tails emit it so; ids grow with the local row): a
            // candidate's merged rank is its own position plus, per other shard, the number of that shard's candidates
            // ahead of it -- one binary search each, no exchange network (8 shards x top-100: 2048-wide bitonic = 66
            // barrier-separated stages over 16-byte records against 8 x 8 probes per candidate)
            int local = 0;
            for (int i = threadIdx.x; i < M; i += blockDim.x) {
                const b200rag_cand e = stage[i];
                if (!e.valid) continue;
                ++local;
                const int sh = i / L;
                int rank = i - sh * L;
                for (int b = 0; b < n_shards && rank < L; ++b) {
                    if (b == sh) continue;
                    const b200rag_cand* lst = stage + (size_t)b * L;
                    int lo = 0, hi = L;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        const b200rag_cand o = lst[mid];
                        if (o.valid && cand_better(o, e)) lo = mid + 1; else hi = mid;
                    }
                    rank += lo;
                }
                if (rank < L) { leg_id[leg * L + rank] = e.id; leg_score[leg * L + rank] = e.score; }
            }
            if (local) atomicAdd(&leg_n[leg], local);
        } else {
            for (int k = 2; k <= Mp; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < Mp; i += blockDim.x) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const b200rag_cand a = stage[i], b = stage[ixj];
                            // "a before b": valid before invalid, then the leg order
                            const bool a_first = a.valid && (!b.valid || cand_better(a, b));
                            const bool b_first = b.valid && (!a.valid || cand_better(b, a));
                            const bool up = (i & k) == 0;           // ascending position = better first
                            if (up ? b_first : a_first) { stage[i] = b; stage[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
            int local = 0;
            for (int i = threadIdx.x; i < Mp; i += blockDim.x) {
                const b200rag_cand e = stage[i];
                if (!e.valid) continue;
                ++local;
                if (i < L) { leg_id[leg * L + i] = e.id; leg_score[leg * L + i] = e.score; }
            }
            if (local) atomicAdd(&leg_n[leg], local);
        }
    }
    __syncthreads();
    const int nd = min(leg_n[0], L);

    if (nlegs == 1) {
        const int n = min(nd, top_k);
        for (int i = threadIdx.x; i < top_k; i += blockDim.x) {
            out_ids[(size_t)q * top_k + i] = i < n ? leg_id[i] : -1;
            out_scores[(size_t)q * top_k + i] = i < n ? (double)leg_score[i] : 0.0;
        }
        if (threadIdx.x == 0) out_counts[q] = n;
        return;
    }

    // Reciprocal rank fusion, qdrant constants: score = sum over legs of 1/(rrf_k + pos0)
    const int ns = min(leg_n[1], L);
    const int NE = nd + ns;
    for (int e = threadIdx.x; e < NE; e += blockDim.x) {
        if (e < nd) {
            const int64_t id = leg_id[e];
            double sc = 1.0 / (double)(rrf_k + e);
            for (int j = 0; j < ns; ++j)
                if (leg_id[L + j] == id) { sc = sc + 1.0 / (double)(rrf_k + j); break; }
            fused[e] = sc; fid[e] = id; ford[e] = e;
        } else {
            const int j = e - nd;
            const int64_t id = leg_id[L + j];
            bool dup = false;
            for (int i = 0; i < nd; ++i)
                if (leg_id[i] == id) { dup = true; break; }
            fused[e] = dup ? -1.0 : 1.0 / (double)(rrf_k + j);
            fid[e] = id; ford[e] = dup ? -1 : e;
        }
        rnk[e] = 0;
    }
    __syncthreads();
    // rank by counting (fused score desc, then first-leg order), P threads per entry so that a wide CTA is busy
    // (top-100: 400 entries on 1024 threads -- 200 compares per thread instead of 2 rounds of 400)
    const int P = NE > 0 ? min(32, max(1, (int)blockDim.x / NE)) : 1;
    for (int it = threadIdx.x; it < NE * P; it += blockDim.x) {
        const int e = it / P, part = it - e * P;
        const int oe = ford[e];
        if (oe < 0) continue;
        const int f0 = (int)((long long)NE * part / P), f1 = (int)((long long)NE * (part + 1) / P);
        const double se = fused[e];
        int c = 0;
        for (int f = f0; f < f1; ++f) {
            const int of = ford[f];
            if (of < 0) continue;
            const double sf = fused[f];
            if (sf > se || (sf == se && of < oe)) ++c;
        }
        if (c) atomicAdd(&rnk[e], c);
    }
    __syncthreads();
    int local = 0;
    for (int e = threadIdx.x; e < NE; e += blockDim.x) {
        if (ford[e] < 0) continue;
        ++local;
        const int rank = rnk[e];
        if (rank < top_k) {
            out_ids[(size_t)q * top_k + rank] = fid[e];
            out_scores[(size_t)q * top_k + rank] = fused[e];
        }
    }
    if (local) atomicAdd(&total_s, local);
    __syncthreads();
    const int n = min(total_s, top_k);
    for (int i = n + threadIdx.x; i < top_k; i += blockDim.x) {
        out_ids[(size_t)q * top_k + i] = -1;
        out_scores[(size_t)q * top_k + i] = 0.0;
    }
    if (threadIdx.x == 0) out_counts[q] = n;
}

// every block p stores this rank's candidate block into slot `rank` of rank p's window, then publishes the epoch
__global__ void __launch_bounds__(256) exchange_kernel(const uint4* __restrict__ mine, int64_t n16,
                                                       void* const* __restrict__ peer_windows, int world, int rank,
                                                       int64_t slot_bytes, int parity, unsigned long long epoch) {
    const int pr = blockIdx.x;
    uint8_t* win = reinterpret_cast<uint8_t*>(peer_windows[pr]);
    uint4* dst = reinterpret_cast<uint4*>(win + ((size_t)parity * world + rank) * slot_bytes);
    for (int64_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = mine[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(win + (size_t)2 * world * slot_bytes) +
                                   (size_t)rank * kFlagStrideU64;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
    }
}

int launch_exchange(Shard* s, const void* mine, int64_t nbytes, void* const* peer_windows_dev, int world, int rank,
                    int64_t slot_bytes, int parity, unsigned long long epoch) {
    // same shared-memory carve-out as the scan kernels: an SM cannot host CTAs of kernels with different carve-outs at
    // the same time, and with the pipelined tail this kernel must co-reside with the next search's scan
    static AttrCache attr;
    if (attr.raise(s->cfg.device, 1))
        B2_CUDA(cudaFuncSetAttribute(exchange_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    exchange_kernel<<<world, 256, 0, s->stream>>>(reinterpret_cast<const uint4*>(mine), nbytes / 16, peer_windows_dev, world,
                                                  rank, slot_bytes, parity, epoch);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

int launch_fuse(Shard* s, int mode, int batch, int L, int top_k, int rrf_k, const b200rag_cand* gathered,
                int n_shards, int has_trailer, int64_t* out_ids, double* out_scores, int32_t* out_counts,
                int64_t shard_stride_override, const unsigned long long* wait_flags, unsigned long long wait_epoch) {
    const int nlegs = mode == B200RAG_HYBRID ? 2 : 1;
    const size_t M = (size_t)next_pow2(n_shards * L);
    const size_t smem = M * sizeof(b200rag_cand) + (size_t)2 * L * (8 + 4 + 8 + 8 + 4 + 4) + 64;
    if (smem > 200 * 1024) { set_error("fuse: n_shards * L too large"); return B200RAG_ERR_INVALID; }
    static AttrCache attr, carve;
    if (smem > 48 * 1024 && attr.raise(s->cfg.device, smem))
        B2_CUDA(cudaFuncSetAttribute(fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (carve.raise(s->cfg.device, 1))     // co-resides with the next search's scan (see exchange_kernel)
        B2_CUDA(cudaFuncSetAttribute(fuse_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int64_t shard_stride = shard_stride_override > 0 ? shard_stride_override
                                                           : (int64_t)nlegs * batch * L + (has_trailer ? 1 : 0);
    // large sets (the merge over many shards, the rank-by-counting of 2L fused entries) want more lanes; after pipelined
    // legs the kernel runs on the second stream and must fit beside the next search's dense-scan CTA
    const bool beside_scan = s->pipeline && !s->pipeline_paused && !s->legs_classic;
    const int fuse_threads = ((size_t)n_shards * L > 512 || L > 64) ? (beside_scan ? 512 : 1024) : 256;
    fuse_kernel<<<batch, fuse_threads, smem, s->stream>>>(gathered, n_shards, shard_stride, has_trailer, nlegs, batch, L, top_k,
                                                 rrf_k, out_ids, out_scores, out_counts, wait_flags, wait_epoch,
                                                 s->x_timeout_cycles);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

}  // namespace b200rag
