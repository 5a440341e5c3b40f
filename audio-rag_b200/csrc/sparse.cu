// K2  sparse_scan: impact-scored inverted index ("BM25" leg) + its builder.
//
// Replaces the sparse leg of client.query_points (reference call sites src/audio_rag/retrieval/qdrant.py:289-293,
// 306-312; arithmetic in qdrant-client local/sparse_distances.py::sparse_dot_product, a Python loop over every
// document).  score[d] = sum_{t in q and d} w_q[t] * w_d[t]; documents sharing no term with the query are
// excluded; BM25 is the special case w_d = tf-idf impact, w_q = query tf.
//
// Index layout (block-major CSR): the shard's documents are cut into blocks of R consecutive local rows.
// Block b owns   dir[b][0..V]   u32 offsets (a dense term directory: one lookup, no search)
//                post_doc[]     u16 byte offset (4 * slot) of the document's accumulator   } SoA, 6 bytes per posting,
//                post_w[]       f32 impact weight                                           } term-major, doc ascending
// Appending rows only ever touches the trailing block, so the index is incremental by construction.
//
// Scan kernel: one CTA per (query, group of consecutive blocks), grid query-fastest.  int32 fixed-point accumulators
// for a block's R documents live in shared memory; all query terms' posting vectors form ONE flat work list (no
// barrier between terms): 128-bit offset loads + 256-bit weight loads, then native shared-memory integer atomics.
// The epilogue applies the eligibility bitmask and keeps a running top-Lc across the CTA's blocks: one pass over the
// accumulators against a threshold that is shared grid-wide per query.  Nothing but Lc keys per CTA reaches HBM.
//
// Roofline: HBM for the algorithmic bytes = sum over query terms of df_shard(t) * 6 (SURVEY 8d counts 4+sizeof(w) = 8);
// measured binding unit: the L1/LSU data pipe (atomic wavefronts), see profiles/README.md.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "engine.h"

namespace b200rag {

constexpr int kSelCapMax = 1024;   // largest Lc; a CTA's running candidate list holds next_pow2(2 Lc) >= 512 keys

struct SparseScanParams {
    const uint32_t* dir;
    const int64_t* blk_base;
    const uint16_t* post_doc;
    const float* post_w;
    int vocab, R;
    int64_t n_rows;
    const int64_t* q_indptr;
    const uint32_t* q_terms;
    const float* q_w;
    const uint32_t* const* masks;  // device [batch] or nullptr
    uint64_t* out;                 // [batch][n_groups][Lc]
    int Lc, n_blocks, sel_cap;
    int bpc, n_groups;             // blocks per CTA, CTAs per query
    int* gthr;                     // [batch] grid-wide pruning thresholds in fixed-point units (INT_MIN = none yet)
    unsigned long long* gkey;      // [batch] the same as a candidate KEY (0 = none yet): the leg tail's survivor threshold
    float w_absmax;                // max |w_d| over the shard's postings: bounds every score by sum|w_q| * w_absmax
    float* q_eps;                  // [batch] out: absolute error bound of this query's approximate scores
    unsigned long long* post_count;
};

// Fixed-point accumulation.  fp32 atomicAdd on shared memory is a CAS loop on sm_100 (ATOMS.CAST.SPIN); the integer
// add is native (ATOMS.ADD).  Each posting therefore adds  (round(w_q*w_d*S) << cb) + 1  to its document's int32
// accumulator: the low `cb` bits count the postings that touched the document (SURVEY R7 needs "touched", not
// "non-zero"), the high bits hold the score in units of 1/S.  S is the largest power of two for which the bound
// sum|w_q| * w_absmax cannot overflow 30-cb bits.  Integer adds commute, so the accumulation needs no ordering and no
// barriers between terms, and the approximate score of a document is the same whatever the schedule.
// The rounding is done by the FFMA itself: w_q*S*w_d + 1.5*2^23 leaves the rounded integer in the low mantissa bits
// (|value| < 2^22), and one IMAD turns those bits into the accumulator increment.
// |approx - exact| <= nterms * (0.5/S + fp32 product rounding); finalize_leg's slack guard gets 1.0*nterms/S.
__device__ __forceinline__ void sparse_scale(float qabs, float w_absmax, int nterms, int* cb_out, float* S_out) {
    const int cb = 32 - __clz(nterms);   // bits that hold a count of up to nterms
    const float bound = qabs * w_absmax;
    int e = 0;
    if (bound > 0.f && bound < 3.0e38f) {
        const int bits = 30 - cb < 22 ? 30 - cb : 22;   // <= 22: round-to-int by adding 1.5 * 2^23 (one FFMA, no F2I)
        e = bits - ilogbf(bound) - 1;
        e = e > 60 ? 60 : (e < -60 ? -60 : e);
    }
    *cb_out = cb;
    *S_out = ldexpf(1.0f, e);
}

// Accumulator slot of a document inside its block.  The postings of a frequent term are (nearly) consecutive
// documents; a thread owns 8 consecutive postings, so at each step the 32 lanes of a warp would hit documents 8 apart:
// 4 distinct banks, an 8-way conflict on every shared-memory atomic (ncu: 4.4 wavefronts per ATOMS).  Folding bits 5..9
// into the bank bits makes every power-of-two stride up to 32 conflict-free.  The map is an involution that keeps
// bits >= 5, so a slot's eligibility word is still word (slot >> 5).  post_doc stores 4 * SLOT, the byte offset of the
// accumulator (applied at build time; docs_per_block <= 16384 keeps it in 16 bits).
__host__ __device__ __forceinline__ uint32_t swz_doc(uint32_t d) { return d ^ ((d >> 5) & 31u); }

// One CTA scans `bpc` consecutive blocks for one query and keeps a running candidate list across them, so its
// pruning threshold strengthens block after block (and is shared grid-wide through gthr[q], like the dense scan):
// in steady state a block's selection is ONE pass over the accumulators with almost nothing pushed.
// 256-bit global load (sm_100: LDG.E.256); p must be 32-byte aligned
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

template <int NT, int EPT, int U, bool MASKED>
__global__ void __launch_bounds__(NT, NT == 128 ? 8 : 4) sparse_scan_kernel(const SparseScanParams p) {
    extern __shared__ __align__(16) uint8_t ssm[];
    constexpr int R = EPT * NT;
    constexpr int kChunk = NT;      // query terms per pass: one term per thread
    constexpr int NW = NT / 32;
    constexpr int kNone = INT_MIN;
    const int cap = p.sel_cap;
    int* acc = reinterpret_cast<int*>(ssm);
    uint64_t* sel = reinterpret_cast<uint64_t*>(ssm + (size_t)R * 4);   // [cap] running candidates of this CTA
    uint32_t* seg_s = reinterpret_cast<uint32_t*>(sel + cap);           // [chunk] first posting of the segment (block-relative)
    uint32_t* seg_e = seg_s + kChunk;                                   // [chunk] one past the last posting
    uint32_t* pref = seg_e + kChunk;                                    // [chunk + 1] vectors before segment j
    float* qw = reinterpret_cast<float*>(pref + kChunk + 1);            // [chunk] w_q * S
    uint32_t* mw_s = reinterpret_cast<uint32_t*>(qw + kChunk);          // [R/32] eligibility words of the block
    __shared__ int cnt_s, c_s, tau_s, wk_s[NW];
    __shared__ uint32_t wsum[NW];
    __shared__ float qabs_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x, g = blockIdx.y;
    const int b_begin = g * p.bpc, b_end = min(p.n_blocks, b_begin + p.bpc);
    const int64_t qs = p.q_indptr[q], qe = p.q_indptr[q + 1];
    const int nterms = (int)(qe - qs);
    uint64_t* out = p.out + ((size_t)q * p.n_groups + g) * p.Lc;

    {
        int4* a4 = reinterpret_cast<int4*>(acc);
        for (int i = tid; i < R / 4; i += NT) a4[i] = make_int4(0, 0, 0, 0);
    }
    if (tid == 0) cnt_s = 0;
    // per-query scale: sum |w_q| in a fixed order (warp 0: strided partials, butterfly)
    if (tid < 32) {
        float a = 0.f;
        for (int64_t i = qs + lane; i < qe; i += 32) a += fabsf(p.q_w[i]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) qabs_s = a;
    }
    __syncthreads();
    int cb;
    float S;
    sparse_scale(qabs_s, p.w_absmax, nterms, &cb, &S);
    if (g == 0 && tid == 0 && p.q_eps != nullptr) p.q_eps[q] = (float)nterms / S + 2.4e-7f * qabs_s * p.w_absmax;
    const float invS = 1.0f / S;
    uint32_t ucmul = 1u << cb;
    const int cmask = (int)ucmul - 1;
    constexpr float kMagic = 12582912.0f;                  // 1.5 * 2^23, bit pattern 0x4B400000
    uint32_t ucadd = 1u - 0x4B400000u * ucmul;             // (bits - 0x4B400000) * 2^cb + 1 == bits * 2^cb + ucadd (mod 2^32)
    asm volatile("" : "+r"(ucmul), "+r"(ucadd));           // opaque: one IMAD per posting instead of VIADD + SHF + VIADD

    // single-chunk queries (the common case) keep their term and weight in registers across blocks and prefetch the
    // next block's directory entries while the current block streams
    const bool one_chunk = nterms <= kChunk;
    uint32_t my_t = 0;
    float my_w = 0.f;
    uint32_t nxt_s = 0, nxt_e = 0;
    if (one_chunk && tid < nterms) {
        my_t = p.q_terms[qs + tid];
        my_w = p.q_w[qs + tid] * S;
        if (b_begin < b_end) {
            const uint32_t* D = p.dir + (size_t)b_begin * (p.vocab + 1);
            nxt_s = D[my_t]; nxt_e = D[my_t + 1];
        }
    }
    unsigned long long npost = 0;   // postings this thread's terms contributed (statistics)
    int tau = kNone;                // pruning threshold in fixed-point score units (documents with v < tau cannot make it)

    // slot -> fixed-point score, or kNone when the document is untouched / not eligible
    auto val_of = [&](int i) -> int {
        const int idx = i * NT + tid;     // a warp reads 32 consecutive slots: one eligibility word per warp per i
        const int a = acc[idx];
        bool ok = (a & cmask) != 0;
        if (MASKED) ok = ok && ((mw_s[idx >> 5] >> (swz_doc((uint32_t)idx) & 31u)) & 1u);
        return ok ? (a >> cb) : kNone;
    };
    auto key_from = [&](int v, int i, int b) -> uint64_t {
        return make_key((float)v * invS + 0.0f, (uint32_t)b * (uint32_t)R + swz_doc((uint32_t)(i * NT + tid)));
    };
    // one pass: append every document with v >= t to the running list; returns false if the list overflowed
    auto push_pass = [&](int t, int b, int base_cnt) -> bool {
#pragma unroll 8
        for (int i = 0; i < EPT; ++i) {
            const int v = val_of(i);
            if (v != kNone && v >= t) {
                const int pos = atomicAdd(&cnt_s, 1);
                if (pos < cap) sel[pos] = key_from(v, i, b);
            }
        }
        __syncthreads();
        const bool ok = cnt_s <= cap;
        __syncthreads();
        if (!ok && tid == 0) cnt_s = base_cnt;
        if (!ok) __syncthreads();
        return ok;
    };
    // sort the running list, keep the best Lc, raise tau to the Lc-th best and publish it
    auto compact = [&]() {
        const int M = cnt_s;
        int npow2 = next_pow2(M > 1 ? M : 1);
        __syncthreads();
        for (int i = M + tid; i < npow2; i += NT) sel[i] = 0;
        cta_bitonic_desc(sel, npow2, tid, NT, 0);
        if (M >= p.Lc) {
            const int t = __float2int_rn(key_score(sel[p.Lc - 1]) * S);   // exact: |v| < 2^23 and S is a power of two
            if (t > tau) tau = t;
            if (tid == 0) {
                cnt_s = p.Lc;
                if (p.gthr != nullptr) atomicMax(&p.gthr[q], tau);
                if (p.gkey != nullptr) atomicMax(&p.gkey[q], (unsigned long long)sel[p.Lc - 1]);   // Lc listed keys are >= it
            }
        }
        __syncthreads();
    };

    for (int b = b_begin; b < b_end; ++b) {
        const uint32_t* D = p.dir + (size_t)b * (p.vocab + 1);
        const int64_t base = p.blk_base[b];      // multiple of 8: posting vectors of the block start at base >> 3
        const uint4* pd = reinterpret_cast<const uint4*>(p.post_doc) + (base >> 3);
        const float4* pw = reinterpret_cast<const float4*>(p.post_w) + (base >> 2);
        bool any = false;
        if (MASKED) {
            const uint32_t* m = p.masks[q];
            for (int i = tid; i < R / 32; i += NT) mw_s[i] = m != nullptr ? m[(size_t)b * (R / 32) + i] : 0xFFFFFFFFu;
        }
        for (int64_t c0 = qs; c0 < qe; c0 += kChunk) {
            const int cn = (int)min((int64_t)kChunk, qe - c0);
            __syncthreads();   // the previous chunk's segments are no longer read; accumulators are zeroed
            // one term per thread: directory lookup, vector count, block-wide exclusive scan
            uint32_t s = 0, e = 0, nv = 0;
            if (tid < cn) {
                float w;
                if (one_chunk) {
                    s = nxt_s; e = nxt_e; w = my_w;
                    if (b + 1 < b_end) { const uint32_t* Dn = D + (p.vocab + 1); nxt_s = Dn[my_t]; nxt_e = Dn[my_t + 1]; }
                } else {
                    const uint32_t t = p.q_terms[c0 + tid];
                    s = D[t]; e = D[t + 1];
                    w = p.q_w[c0 + tid] * S;
                }
                if (e > s) { nv = ((e + 7) >> 3) - (s >> 3); npost += e - s; }
                seg_s[tid] = s; seg_e[tid] = e; qw[tid] = w;
            }
            uint32_t incl = nv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            uint32_t woff = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) woff += (w < warp) ? wsum[w] : 0u;
            if (tid < cn) pref[tid] = woff + incl - nv;
            if (tid == cn - 1) pref[cn] = woff + incl;
            __syncthreads();
            const uint32_t total = pref[cn];
            any = any || total != 0;

            // flat work list over the vectors of all the chunk's segments: U independent 48-byte loads per thread in
            // flight, then 8*U native shared-memory integer atomics.  No barrier between terms.
            int j = 0;
            for (uint32_t v0 = tid; v0 < total; v0 += NT * U) {
                uint4 d4[U];
                float4 wa[U], wb[U];
                uint32_t lo[U], hi[U], P[U];
                float wq[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t v = v0 + u * NT;
                    lo[u] = 1; hi[u] = 0; P[u] = 0; wq[u] = 0.f;
                    if (v < total) {
                        while (v >= pref[j + 1]) ++j;
                        const uint32_t ss = seg_s[j];
                        const uint32_t vec = (ss >> 3) + (v - pref[j]);
                        d4[u] = pd[vec];
                        ldg256(pw + 2 * (size_t)vec, wa[u], wb[u]);   // one 32-byte load: 8 wavefronts per warp instead of 2 x 8
                        lo[u] = ss; hi[u] = seg_e[j]; P[u] = vec << 3; wq[u] = qw[j];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (lo[u] >= hi[u]) continue;
                    const uint32_t dd[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
                    const float ww[8] = {wa[u].x, wa[u].y, wa[u].z, wa[u].w, wb[u].x, wb[u].y, wb[u].z, wb[u].w};
                    if (P[u] >= lo[u] && P[u] + 8 <= hi[u]) {   // interior vector: all 8 postings belong to the segment
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t off = (k & 1) ? (dd[k >> 1] >> 16) : (dd[k >> 1] & 0xFFFFu);   // 4 * slot
                            atomicAdd(reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(acc) + off),
                                      (int)(__float_as_uint(fmaf(wq[u], ww[k], kMagic)) * ucmul + ucadd));
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t Pk = P[u] + k;
                            if (Pk >= lo[u] && Pk < hi[u]) {
                                const uint32_t off = (k & 1) ? (dd[k >> 1] >> 16) : (dd[k >> 1] & 0xFFFFu);
                                atomicAdd(reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(acc) + off),
                                          (int)(__float_as_uint(fmaf(wq[u], ww[k], kMagic)) * ucmul + ucadd));
                            }
                        }
                    }
                }
            }
        }
        if (tid == 0) tau_s = p.gthr != nullptr ? *reinterpret_cast<volatile int*>(&p.gthr[q]) : kNone;   // fresh every block
        __syncthreads();
        if (!any) continue;   // block shares no term with the query: accumulators are still zero (uniform branch)
        if (tau_s > (int)0x80808080 && tau_s > tau) tau = tau_s;   // (memset pattern 0x80808080 == none yet)

        // ---- selection
        if (cnt_s > cap - p.Lc) compact();   // (cap >= 2 Lc: after this at least Lc slots are free)
        const int base_cnt = cnt_s;
        __syncthreads();
        bool done = false;
        if (tau != kNone) done = push_pass(tau, b, base_cnt);
        if (!done) {
            // No threshold yet (first block), or the list overflowed.  Block-local bound: every warp sorts its 32 thread
            // maxima in registers and publishes its k-th largest, k = ceil(Lc / #warps); the minimum over the warps has
            // >= Lc distinct documents at or above it (or every touched document is at or above it).
            const int kk = (p.Lc + NW - 1) / NW;
            if (kk <= 32) {
                int tmax = kNone;
#pragma unroll 8
                for (int i = 0; i < EPT; ++i) tmax = max(tmax, val_of(i));
                int v = tmax;
#pragma unroll
                for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        const int o = __shfl_xor_sync(0xffffffffu, v, j);
                        const bool keep_max = (((lane & j) == 0) == ((lane & k) == 0));
                        v = keep_max ? max(o, v) : min(o, v);
                    }
                const int kth = __shfl_sync(0xffffffffu, v, kk - 1);
                if (lane == 0) wk_s[warp] = kth;
            } else {
                // large Lc: every thread finds its T-th best document, T = ceil(Lc / #threads); the minimum over the
                // threads has >= T documents per thread at or above it
                const int T = (p.Lc + NT - 1) / NT;
                int prev_v = INT_MAX, prev_i = -1, best_v = kNone;
                for (int r = 0; r < T; ++r) {
                    int best_i = -1;
                    best_v = kNone;
                    for (int i = 0; i < EPT; ++i) {
                        const int v = val_of(i);
                        const bool after = v < prev_v || (v == prev_v && i > prev_i);
                        if (v != kNone && after && v > best_v) { best_v = v; best_i = i; }
                    }
                    if (best_v == kNone) break;
                    prev_v = best_v; prev_i = best_i;
                }
                int v = best_v;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
                if (lane == 0) wk_s[warp] = v;
            }
            __syncthreads();
            int t2 = wk_s[0];
#pragma unroll
            for (int w = 1; w < NW; ++w) t2 = min(t2, wk_s[w]);
            if (tau > t2) t2 = tau;
            __syncthreads();
            done = push_pass(t2, b, base_cnt);
        }
        if (!done) {
            // Rare (massive ties, or Lc > 32 * #warps): find the block's exact Lc-th largest KEY by bisection on the key
            // bits (keys are unique), then append exactly the keys >= it.
            uint64_t K = 0;
            for (int bit = 63; bit >= 0; --bit) {
                const uint64_t cand = K | (1ull << bit);
                int c = 0;
                for (int i = 0; i < EPT; ++i) {
                    const int v = val_of(i);
                    c += (v != kNone && key_from(v, i, b) >= cand) ? 1 : 0;
                }
                if (tid == 0) c_s = 0;
                __syncthreads();
                c = __reduce_add_sync(0xffffffffu, c);
                if (lane == 0 && c) atomicAdd(&c_s, c);
                __syncthreads();
                if (c_s >= p.Lc) K = cand;
                __syncthreads();
            }
            for (int i = 0; i < EPT; ++i) {
                const int v = val_of(i);
                if (v == kNone) continue;
                const uint64_t k = key_from(v, i, b);
                if (k >= K) {
                    const int pos = atomicAdd(&cnt_s, 1);
                    if (pos < cap) sel[pos] = k;
                }
            }
            __syncthreads();
        }
        if (cnt_s >= 2 * p.Lc) compact();
        // zero the accumulators for the next block
        if (b + 1 < b_end) {
            int4* a4 = reinterpret_cast<int4*>(acc);
            for (int i = tid; i < R / 4; i += NT) a4[i] = make_int4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    compact();   // sorted, at most Lc entries count
    const int M = min(cnt_s, p.Lc);
    for (int i = tid; i < p.Lc; i += NT) out[i] = i < M ? sel[i] : 0ull;
    if (p.post_count != nullptr) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) npost += __shfl_xor_sync(0xffffffffu, npost, d);
        if (lane == 0 && npost) atomicAdd(p.post_count, npost);
    }
}

template <int NT, int EPT, int U>
static int launch_scan_tu(Shard* s, const SparseScanParams& p, int batch) {
    const size_t smem = (size_t)EPT * NT * 4 + (size_t)p.sel_cap * 8 + (size_t)NT * 16 + 16 + (size_t)EPT * NT / 8;
    auto kern = p.masks != nullptr ? sparse_scan_kernel<NT, EPT, U, true> : sparse_scan_kernel<NT, EPT, U, false>;
    static AttrCache attr[2];                  // per instantiation and mask flavour
    if (attr[p.masks != nullptr ? 1 : 0].raise(s->cfg.device, smem)) {
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    dim3 grid((unsigned)batch, (unsigned)p.n_groups);
    kern<<<grid, NT, smem, s->stream>>>(p);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

template <int NT>
static int launch_scan_nt(Shard* s, const SparseScanParams& p, int batch) {
    switch (s->R / NT) {
        case 8: return launch_scan_tu<NT, 8, 2>(s, p, batch);
        case 16: return launch_scan_tu<NT, 16, 2>(s, p, batch);
        case 32: return launch_scan_tu<NT, 32, 2>(s, p, batch);
        case 64: return launch_scan_tu<NT, 64, 2>(s, p, batch);
        case 128: return launch_scan_tu<NT, 128, 2>(s, p, batch);
        default: break;
    }
    set_error("sparse_scan: unsupported docs_per_block");
    return B200RAG_ERR_INVALID;
}

// blocks per CTA: as many as possible (the running threshold prunes better) while the grid still has ~4 waves of CTAs
int sparse_scan_bpc(const Shard* s, int batch, int Lc) {
    (void)Lc;      // (the fused leg tail takes up to 2^20 keys: 12.5M rows x top-100 = 1 526 lists x 300 keys fit as they are)
    if (s->sparse_bpc > 0) return s->sparse_bpc;
    const int64_t items = (int64_t)s->n_blocks * batch;
    int64_t bpc = items / ((int64_t)4 * s->sm_count * 8);
    if (bpc < 1) bpc = 1;
    if (bpc > 32) bpc = 32;
    return (int)bpc;
}
int sparse_scan_nlists(const Shard* s, int batch, int Lc) {
    const int bpc = sparse_scan_bpc(s, batch, Lc);
    return (int)((s->n_blocks + bpc - 1) / bpc);
}

int launch_sparse_scan(Shard* s, int batch, int Lc, uint64_t* out_lists, float* q_eps, int* gthr, uint64_t* gkey) {
    SparseScanParams p{};
    p.dir = s->dir.as<uint32_t>();
    p.blk_base = s->blk_base.as<int64_t>();
    p.post_doc = s->post_doc.as<uint16_t>();
    p.post_w = s->post_w.as<float>();
    p.vocab = s->vocab;
    p.R = s->R;
    p.n_rows = s->n_rows;
    p.q_indptr = s->ws.q_sp_indptr.as<int64_t>();
    p.q_terms = s->ws.q_sp_terms.as<uint32_t>();
    p.q_w = s->ws.q_sp_w.as<float>();
    p.masks = s->h_masks.empty() ? nullptr : s->ws.q_masks.as<const uint32_t*>();
    p.out = out_lists;
    p.Lc = Lc;
    p.n_blocks = (int)s->n_blocks;
    p.bpc = sparse_scan_bpc(s, batch, Lc);
    p.n_groups = sparse_scan_nlists(s, batch, Lc);
    p.sel_cap = next_pow2(2 * Lc) < 512 ? 512 : next_pow2(2 * Lc);
    p.w_absmax = s->w_absmax;
    p.q_eps = q_eps;
    p.gthr = gthr;
    p.gkey = reinterpret_cast<unsigned long long*>(gkey);
    p.post_count = s->ws.post_count.as<unsigned long long>();
    if (Lc > kSelCapMax) { set_error("sparse_scan: top-k too large"); return B200RAG_ERR_INVALID; }
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[2], s->stream)); }
    // CTA size: 128 threads for single queries (more CTAs per SM hide the per-block latency chain), 256 for batches
    // (measured on B200, 10M rows, B = 64: 2.04 ms vs 2.23 ms with 128 threads; B = 1: 74 vs 76 us, a wash), and 256
    // whenever the block is too large for 128 x 128 accumulators per thread.  B200RAG_SPARSE_THREADS forces either.
    const bool forced = getenv("B200RAG_SPARSE_THREADS") != nullptr;
    const bool want256 = forced ? s->sparse_threads == 256 : batch >= 16;
    const int nt = ((want256 && s->R / 256 >= 8) || s->R / 128 > 128 || s->R / 128 < 8) ? 256 : 128;
    if (nt == 256 && s->R / 256 < 8) { set_error("sparse_scan: docs_per_block too small"); return B200RAG_ERR_INVALID; }
    const int rc = nt == 128 ? launch_scan_nt<128>(s, p, batch) : launch_scan_nt<256>(s, p, batch);
    if (rc == B200RAG_OK && s->profile) { B2_CUDA(cudaEventRecord(s->ev[3], s->stream)); s->ev_sparse = true; }
    return rc;
}

// ================================================================================================ builder
__global__ void gather_i64_kernel(const int64_t* __restrict__ src, int64_t stride, int64_t n_last, int64_t count,
                                  int64_t* __restrict__ dst) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < count) {
        int64_t idx = i * stride;
        if (idx > n_last) idx = n_last;
        dst[i] = src[idx];
    }
}

// vals[i - p0] = (doc_in_block << 32) | weight bits, one warp per document
__global__ void pack_vals_kernel(const int64_t* __restrict__ fwd_ptr, const float* __restrict__ fwd_w, int64_t d0,
                                 int64_t d1, int64_t p0, uint64_t* __restrict__ vals) {
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t d = d0 + w;
    if (d >= d1) return;
    const int64_t s = fwd_ptr[d], e = fwd_ptr[d + 1];
    for (int64_t i = s + lane; i < e; i += 32)
        vals[i - p0] = ((uint64_t)(uint32_t)(d - d0) << 32) | (uint64_t)__float_as_uint(fwd_w[i]);
}

__global__ void unpack_block_kernel(const uint64_t* __restrict__ vals, int64_t cnt, uint16_t* __restrict__ post_doc,
                                    float* __restrict__ post_w, int64_t base, int64_t padded) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < cnt) {
        const uint64_t v = vals[i];
        post_doc[base + i] = (uint16_t)(4u * swz_doc((uint32_t)(v >> 32)));   // byte offset of the accumulator slot
        post_w[base + i] = __uint_as_float((uint32_t)v);
    } else if (i < padded) {
        post_doc[base + i] = 0;
        post_w[base + i] = 0.f;
    }
}

// dir[t] = number of postings of the block with term < t  (lower bound in the sorted term keys), t in [0, V]
__global__ void directory_kernel(const uint32_t* __restrict__ keys, int64_t cnt, int vocab, uint32_t* __restrict__ dir) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t > vocab) return;
    int64_t lo = 0, hi = cnt;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)keys[mid] < t) lo = mid + 1; else hi = mid;
    }
    dir[t] = (uint32_t)lo;
}

__global__ void absmax_kernel(const float* __restrict__ w, int64_t n, uint32_t* __restrict__ out_bits) {
    uint32_t m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t u = __float_as_uint(fabsf(w[i]));   // non-negative floats order like their bit patterns
        m = u > m ? u : m;
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out_bits, m);
}

// max |w_d| over the postings appended since the last call (the sparse scan's fixed-point scale needs the bound)
static int update_w_absmax(Shard* s) {
    if (s->wmax_nnz == s->nnz) return B200RAG_OK;
    cudaStream_t st = s->stream;
    DevBuf d;
    B2_TRY(d.ensure(4, 0, st));
    B2_CUDA(cudaMemsetAsync(d.p, 0, 4, st));
    const int64_t n = s->nnz - s->wmax_nnz;
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)s->sm_count * 8);
    absmax_kernel<<<grid, 256, 0, st>>>(s->fwd_w.as<float>() + s->wmax_nnz, n, d.as<uint32_t>());
    uint32_t bits = 0;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bits, d.p, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    d.release();
    if (e != cudaSuccess) return cuda_fail(e, "absmax");
    float f;
    memcpy(&f, &bits, 4);
    if (f > s->w_absmax) s->w_absmax = f;
    s->wmax_nnz = s->nnz;
    return B200RAG_OK;
}

int build_inverted(Shard* s) {
    cudaStream_t st = s->stream;
    B2_TRY(update_w_absmax(s));
    const int R = s->R;
    const int64_t n = s->n_rows;
    const int64_t nb_new = (n + R - 1) / R;
    if (n == s->built_rows) return B200RAG_OK;
    const int64_t b0 = s->built_rows / R;  // blocks below b0 are complete and stay as they are

    // block boundaries of the forward index
    const int64_t nbnd = nb_new + 1;
    DevBuf bnd_d;
    B2_TRY(bnd_d.ensure((size_t)nbnd * 8, 0, st));
    gather_i64_kernel<<<(unsigned)((nbnd + 255) / 256), 256, 0, st>>>(s->fwd_ptr.as<int64_t>(), R, n, nbnd,
                                                                      bnd_d.as<int64_t>());
    B2_CUDA(cudaGetLastError());
    std::vector<int64_t> bnd((size_t)nbnd);
    B2_CUDA(cudaMemcpyAsync(bnd.data(), bnd_d.p, (size_t)nbnd * 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));

    s->h_blk_base.resize((size_t)nb_new + 1);
    if (b0 == 0) s->h_blk_base[0] = 0;
    int64_t max_cnt = 0;
    for (int64_t b = b0; b < nb_new; ++b) {
        const int64_t cnt = bnd[b + 1] - bnd[b];
        if (cnt > max_cnt) max_cnt = cnt;
        if (cnt > 0xFFFFFFF0ll) { set_error("build: block has too many postings"); bnd_d.release(); return B200RAG_ERR_INVALID; }
        s->h_blk_base[b + 1] = s->h_blk_base[b] + ((cnt + 7) & ~7ll);
    }
    const int64_t inv_total = s->h_blk_base[nb_new];
    const int64_t keep = s->h_blk_base[b0];
    B2_TRY(s->post_doc.ensure((size_t)(inv_total + 8) * 2, (size_t)keep * 2, st));
    B2_TRY(s->post_w.ensure((size_t)(inv_total + 8) * 4, (size_t)keep * 4, st));
    B2_TRY(s->dir.ensure((size_t)nb_new * (s->vocab + 1) * 4, (size_t)b0 * (s->vocab + 1) * 4, st));
    B2_TRY(s->blk_base.ensure((size_t)(nb_new + 1) * 8, 0, st));
    B2_CUDA(cudaMemcpyAsync(s->blk_base.p, s->h_blk_base.data(), (size_t)(nb_new + 1) * 8, cudaMemcpyHostToDevice, st));

    DevBuf keys_out, vals_in, vals_out, temp;
    int rc = B200RAG_OK;
    size_t temp_bytes = 0;
    int end_bit = 1;
    while ((1ll << end_bit) < s->vocab) ++end_bit;
    if (max_cnt > 0) {
        cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                        (const uint64_t*)nullptr, (uint64_t*)nullptr, max_cnt, 0, end_bit, st);
        if ((rc = keys_out.ensure((size_t)max_cnt * 4, 0, st)) != B200RAG_OK) goto done;
        if ((rc = vals_in.ensure((size_t)max_cnt * 8, 0, st)) != B200RAG_OK) goto done;
        if ((rc = vals_out.ensure((size_t)max_cnt * 8, 0, st)) != B200RAG_OK) goto done;
        if ((rc = temp.ensure(temp_bytes + 16, 0, st)) != B200RAG_OK) goto done;
    }
    for (int64_t b = b0; b < nb_new; ++b) {
        const int64_t d0 = b * R, d1 = (d0 + R < n) ? d0 + R : n;
        const int64_t p0 = bnd[b], cnt = bnd[b + 1] - bnd[b];
        const int64_t base = s->h_blk_base[b];
        uint32_t* dir_b = s->dir.as<uint32_t>() + (size_t)b * (s->vocab + 1);
        if (cnt > 0) {
            const int64_t nthreads = (d1 - d0) * 32;
            pack_vals_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(
                s->fwd_ptr.as<int64_t>(), s->fwd_w.as<float>(), d0, d1, p0, vals_in.as<uint64_t>());
            size_t tb = temp_bytes;
            cudaError_t e = cub::DeviceRadixSort::SortPairs(temp.p, tb, s->fwd_terms.as<uint32_t>() + p0,
                                                            keys_out.as<uint32_t>(), vals_in.as<uint64_t>(),
                                                            vals_out.as<uint64_t>(), cnt, 0, end_bit, st);
            if (e != cudaSuccess) { rc = cuda_fail(e, "cub::DeviceRadixSort::SortPairs"); goto done; }
            const int64_t padded = (cnt + 7) & ~7ll;
            unpack_block_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, st>>>(
                vals_out.as<uint64_t>(), cnt, s->post_doc.as<uint16_t>(), s->post_w.as<float>(), base, padded);
        }
        directory_kernel<<<(unsigned)((s->vocab + 1 + 255) / 256), 256, 0, st>>>(keys_out.as<uint32_t>(), cnt,
                                                                                s->vocab, dir_b);
    }
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { rc = cuda_fail(e, "build kernels"); goto done; }
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "build sync"); goto done; }
    }
    s->n_blocks = nb_new;
    s->inv_nnz = inv_total;
    s->built_rows = n;
done:
    keys_out.release(); vals_in.release(); vals_out.release(); temp.release(); bnd_d.release();
    return rc;
}

}  // namespace b200rag
