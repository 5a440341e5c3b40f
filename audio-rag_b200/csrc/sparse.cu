// K2  sparse_scan: impact-scored inverted index ("BM25" leg) + its builder.
//
// Replaces the sparse leg of client.query_points (reference call sites src/audio_rag/retrieval/qdrant.py:289-293,
// 306-312; arithmetic in qdrant-client local/sparse_distances.py::sparse_dot_product, a Python loop over every
// document).  score[d] = sum_{t in q and d} w_q[t] * w_d[t]; documents sharing no term with the query are
// excluded; BM25 is the special case w_d = tf-idf impact, w_q = query tf.
//
// Index layout (block-major CSR): the shard's documents are cut into blocks of R consecutive local rows.
// Block b owns   dir[b][0..V]   u32 offsets (a dense term directory: one lookup, no search)
//                post_doc[]     u16 document offset inside the block      } SoA, 6 bytes per posting,
//                post_w[]       f32 impact weight                          } term-major, doc ascending
// Appending rows only ever touches the trailing block, so the index is incremental by construction.
//
// Scan kernel: one CTA per (block, query).  fp32 accumulators for the block's R documents live in shared
// memory (-0.0f == "untouched"); the query's terms are walked in ascending term id, each term's postings
// segment streamed with aligned 128/256-bit loads and applied with plain (conflict-free within a term)
// shared-memory read-modify-writes, so the summation order per document is deterministic.  The epilogue
// applies the eligibility bitmask and selects the block's top-Lc without sorting the block: per-thread maxima
// give a threshold, survivors are compacted and bitonic-sorted.  Nothing but Lc keys per (block, query)
// reaches HBM.
//
// Roofline: HBM.  Algorithmic bytes = sum over query terms of df_shard(t) * 6  (SURVEY 8d counts 4+sizeof(w) = 8).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "engine.h"

namespace b200rag {

constexpr int kSelCap = 1024;

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
    uint64_t* out;                 // [batch][n_blocks][Lc]
    int Lc, n_blocks;
    unsigned long long* post_count;
};

template <int EPT>
__global__ void __launch_bounds__(kSparseThreads) sparse_scan_kernel(const SparseScanParams p) {
    extern __shared__ __align__(16) uint8_t ssm[];
    constexpr int NT = kSparseThreads;
    constexpr int R = EPT * NT;
    float* acc = reinterpret_cast<float*>(ssm);
    uint64_t* sel = reinterpret_cast<uint64_t*>(ssm + (size_t)R * 4);
    uint32_t* seg_s = reinterpret_cast<uint32_t*>(sel + kSelCap);
    uint32_t* seg_e = seg_s + kMaxQueryTermsChunk;
    float* qw = reinterpret_cast<float*>(seg_e + kMaxQueryTermsChunk);
    uint32_t* seg_s_raw = reinterpret_cast<uint32_t*>(qw + kMaxQueryTermsChunk);
    uint32_t* seg_e_raw = seg_s_raw + kMaxQueryTermsChunk;
    float* qw_raw = reinterpret_cast<float*>(seg_e_raw + kMaxQueryTermsChunk);
    uint32_t* mw_s = reinterpret_cast<uint32_t*>(qw_raw + kMaxQueryTermsChunk);   // [R/32] eligibility words of the block
    __shared__ int cnt_s;
    __shared__ unsigned long long npost_s;

    const int tid = threadIdx.x;
    const int b = blockIdx.x, q = blockIdx.y;
    const uint32_t* D = p.dir + (size_t)b * (p.vocab + 1);
    const int64_t base = p.blk_base[b];
    const int64_t qs = p.q_indptr[q], qe = p.q_indptr[q + 1];
    uint64_t* out = p.out + ((size_t)q * p.n_blocks + b) * p.Lc;

    for (int i = tid; i < R; i += NT) acc[i] = -0.0f;
    if (tid == 0) { cnt_s = 0; npost_s = 0; }

    const uint4* pd = reinterpret_cast<const uint4*>(p.post_doc);
    const float4* pw = reinterpret_cast<const float4*>(p.post_w);

    __shared__ int nne_s;
    for (int64_t c0 = qs; c0 < qe; c0 += kMaxQueryTermsChunk) {
        const int cn = (int)min((int64_t)kMaxQueryTermsChunk, qe - c0);
        __syncthreads();
        for (int j = tid; j < cn; j += NT) {
            const uint32_t t = p.q_terms[c0 + j];
            qw_raw[j] = p.q_w[c0 + j];
            const uint32_t s = D[t], e = D[t + 1];
            seg_s_raw[j] = s;
            seg_e_raw[j] = e;
            if (e > s) atomicAdd(&npost_s, (unsigned long long)(e - s));
        }
        __syncthreads();
        // warp 0 compacts the terms that have postings in this block (order preserved: ascending term id)
        if (tid < 32) {
            int base_n = 0;
            for (int j0 = 0; j0 < cn; j0 += 32) {
                const int j = j0 + tid;
                const bool ne = j < cn && seg_e_raw[j] > seg_s_raw[j];
                const unsigned bal = __ballot_sync(0xffffffffu, ne);
                if (ne) {
                    const int pos = base_n + __popc(bal & ((1u << tid) - 1u));
                    seg_s[pos] = seg_s_raw[j];
                    seg_e[pos] = seg_e_raw[j];
                    qw[pos] = qw_raw[j];
                }
                base_n += __popc(bal);
            }
            if (tid == 0) nne_s = base_n;
        }
        __syncthreads();
        const int nne = nne_s;
        // software pipeline: the first vector of term j+1 is in flight while term j is applied
        uint4 nd4 = make_uint4(0, 0, 0, 0);
        float4 nwa = make_float4(0, 0, 0, 0), nwb = nwa;
        if (nne > 0) {
            const int64_t P0 = base + seg_s[0], P1 = base + seg_e[0];
            const int64_t vec = (P0 >> 3) + tid;
            if ((vec << 3) < P1) { nd4 = pd[vec]; nwa = pw[2 * vec]; nwb = pw[2 * vec + 1]; }
        }
        for (int j = 0; j < nne; ++j) {
            const float wq = qw[j];
            const int64_t P0 = base + seg_s[j], P1 = base + seg_e[j];
            uint4 d4 = nd4;
            float4 wa = nwa, wb = nwb;
            if (j + 1 < nne) {
                const int64_t Q0 = base + seg_s[j + 1], Q1 = base + seg_e[j + 1];
                const int64_t nvec = (Q0 >> 3) + tid;
                if ((nvec << 3) < Q1) { nd4 = pd[nvec]; nwa = pw[2 * nvec]; nwb = pw[2 * nvec + 1]; }
            }
            for (int64_t vec = (P0 >> 3) + tid; (vec << 3) < P1; vec += NT) {
                if (vec != (P0 >> 3) + tid) { d4 = pd[vec]; wa = pw[2 * vec]; wb = pw[2 * vec + 1]; }
                const int64_t P = vec << 3;
                const uint32_t dd[4] = {d4.x, d4.y, d4.z, d4.w};
                const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int64_t Pk = P + k;
                    if (Pk >= P0 && Pk < P1) {
                        const uint32_t d = (k & 1) ? (dd[k >> 1] >> 16) : (dd[k >> 1] & 0xFFFFu);
                        acc[d] = __fadd_rn(acc[d], __fadd_rn(__fmul_rn(wq, ww[k]), 0.0f));
                    }
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (npost_s == 0) {  // block shares no term with the query
        for (int i = tid; i < p.Lc; i += NT) out[i] = 0;
        return;
    }
    if (tid == 0 && p.post_count != nullptr) atomicAdd(p.post_count, npost_s);

    // ---- selection.  Keys are recomputed from the accumulators on each pass (keeps registers low -> more CTAs/SM).
    const uint32_t* m = p.masks != nullptr ? p.masks[q] : nullptr;
    // the block's eligibility words go to shared memory once (R/32 words, coalesced); sel[] is free until the push
    const bool use_mask = m != nullptr;
    if (use_mask)
        for (int i = tid; i < R / 32; i += NT) mw_s[i] = m[(size_t)b * (R / 32) + i];
    __syncthreads();
    auto key_of = [&](int i) -> uint64_t {
        const int idx = i * NT + tid;
        const float v = acc[idx];
        const uint32_t doc = (uint32_t)b * (uint32_t)R + (uint32_t)idx;
        bool ok = __float_as_uint(v) != 0x80000000u;
        if (ok && use_mask) ok = (mw_s[idx >> 5] >> (idx & 31)) & 1u;
        return ok ? make_key(v, doc) : 0ull;
    };
    uint64_t tmax = 0;
#pragma unroll 8
    for (int i = 0; i < EPT; ++i) {
        const uint64_t k = key_of(i);
        tmax = k > tmax ? k : tmax;
    }
    // threshold: every warp sorts its 32 thread maxima in registers and publishes its k-th largest, k = ceil(Lc / #warps);
    // the minimum over the warps has >= Lc distinct documents at or above it, so it is a valid lower bound for the
    // block's Lc-th best key (no block-wide sort, one barrier).
    constexpr int NW = NT / 32;
    __shared__ uint64_t wk_s[NW];
    {
        const int lane = tid & 31;
        uint64_t v = tmax;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const uint64_t o = __shfl_xor_sync(0xffffffffu, v, j);
                const bool keep_max = (((lane & j) == 0) == ((lane & k) == 0));
                v = keep_max ? (o > v ? o : v) : (o < v ? o : v);
            }
        const int kk = (p.Lc + NW - 1) / NW;
        const uint64_t kth = kk <= 32 ? __shfl_sync(0xffffffffu, v, kk - 1) : 0ull;
        if (lane == 0) wk_s[tid >> 5] = kth;
    }
    __syncthreads();
    uint64_t tau = wk_s[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) tau = wk_s[w] < tau ? wk_s[w] : tau;
#pragma unroll 8
    for (int i = 0; i < EPT; ++i) {
        const uint64_t k = key_of(i);
        if (k != 0 && k >= tau) {
            const int pos = atomicAdd(&cnt_s, 1);
            if (pos < kSelCap) sel[pos] = k;
        }
    }
    __syncthreads();
    int M = cnt_s;
    if (M > kSelCap) {
        // Rare: more than kSelCap documents at or above the threshold estimate (e.g. massive ties).  Find the exact
        // Lc-th largest key by bisection on the key bits (keys are unique), then collect exactly the keys >= it.
        __shared__ int c_s;
        uint64_t K = 0;
        for (int bit = 63; bit >= 0; --bit) {
            const uint64_t cand = K | (1ull << bit);
            int c = 0;
            for (int i = 0; i < EPT; ++i) c += key_of(i) >= cand ? 1 : 0;
            if (tid == 0) c_s = 0;
            __syncthreads();
            c = __reduce_add_sync(0xffffffffu, c);
            if ((tid & 31) == 0 && c) atomicAdd(&c_s, c);
            __syncthreads();
            if (c_s >= p.Lc) K = cand;
            __syncthreads();
        }
        if (tid == 0) cnt_s = 0;
        __syncthreads();
        for (int i = 0; i < EPT; ++i) {
            const uint64_t k = key_of(i);
            if (k >= K && k != 0) {
                const int pos = atomicAdd(&cnt_s, 1);
                if (pos < kSelCap) sel[pos] = k;
            }
        }
        __syncthreads();
        M = cnt_s;
    }
    int npow2 = next_pow2(M > p.Lc ? M : p.Lc);
    if (npow2 > kSelCap) npow2 = kSelCap;
    for (int i = M + tid; i < npow2; i += NT) sel[i] = 0;
    cta_bitonic_desc(sel, npow2, tid, NT, 0);
    for (int i = tid; i < p.Lc; i += NT) out[i] = i < npow2 ? sel[i] : 0ull;
}

template <int EPT>
static int launch_scan_t(Shard* s, const SparseScanParams& p, int batch) {
    const size_t smem = (size_t)EPT * kSparseThreads * 4 + (size_t)kSelCap * 8 + (size_t)kMaxQueryTermsChunk * 24 +
                        (size_t)EPT * kSparseThreads / 8;
    auto kern = sparse_scan_kernel<EPT>;
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    dim3 grid((unsigned)s->n_blocks, (unsigned)batch);
    kern<<<grid, kSparseThreads, smem, s->stream>>>(p);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

int launch_sparse_scan(Shard* s, int batch, int Lc, uint64_t* out_lists) {
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
    p.post_count = s->ws.post_count.as<unsigned long long>();
    if (Lc > kSelCap) { set_error("sparse_scan: top-k too large"); return B200RAG_ERR_INVALID; }
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[2], s->stream)); }
    int rc;
    switch (s->R / kSparseThreads) {
        case 4: rc = launch_scan_t<4>(s, p, batch); break;
        case 8: rc = launch_scan_t<8>(s, p, batch); break;
        case 16: rc = launch_scan_t<16>(s, p, batch); break;
        case 32: rc = launch_scan_t<32>(s, p, batch); break;
        case 64: rc = launch_scan_t<64>(s, p, batch); break;
        case 128: rc = launch_scan_t<128>(s, p, batch); break;
        default: set_error("sparse_scan: unsupported docs_per_block"); return B200RAG_ERR_INVALID;
    }
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
        post_doc[base + i] = (uint16_t)(v >> 32);
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

int build_inverted(Shard* s) {
    cudaStream_t st = s->stream;
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
