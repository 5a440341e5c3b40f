// Device twins of audio-rag_b200/b200rag/synth.py (bit-identical; integer hash + integer thresholds + IEEE fp64 div).
// Bench / test utilities: they let a 10M-100M row shard be materialised directly in HBM.
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "engine.h"

namespace b200rag {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t stream_key(uint64_t seed, uint64_t stream) {
    return mix64(seed * 0x10000ull + stream);
}
__host__ __device__ __forceinline__ uint64_t row_key(uint64_t skey, uint64_t row) {
    return mix64(skey ^ (row * 0xD6E8FEB86659FD93ull));
}
__device__ __forceinline__ int raw_int(uint64_t rkey, uint64_t j) {
    const uint64_t h = mix64(rkey + j);
    return (int)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + ((h >> 48) & 0xFFFF)) - 131070;
}
__device__ __forceinline__ uint16_t f32_to_bf16_rne(float y) {
    uint32_t u = __float_as_uint(y);
    u = u + 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// one warp per row; lane handles elements lane, lane+32, ...
__global__ void synth_dense_kernel(uint64_t skey, int64_t row0, int64_t n, int dim, uint16_t* __restrict__ out) {
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint64_t rk = row_key(skey, (uint64_t)(row0 + w));
    long long ss = 0;
    for (int k = lane; k < dim; k += 32) {
        const long long x = raw_int(rk, (uint64_t)k);
        ss += x * x;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if (ss == 0) ss = 1;
    const double nrm = sqrt((double)ss);
    for (int k = lane; k < dim; k += 32) {
        const float y = (float)((double)raw_int(rk, (uint64_t)k) / nrm);
        out[w * (int64_t)dim + k] = f32_to_bf16_rne(y);
    }
}

int launch_synth_dense(cudaStream_t st, uint64_t seed, int64_t row0, int64_t n, int dim, uint16_t* out) {
    if (n <= 0) return B200RAG_OK;
    const int64_t threads = n * 32;
    synth_dense_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(stream_key(seed, 1), row0, n, dim, out);
    B2_CUDA(cudaGetLastError());
    return B200RAG_OK;
}

__device__ __forceinline__ int zipf_rank(const uint64_t* __restrict__ thr, int v, uint64_t u) {
    int lo = 0, hi = v;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (thr[mid] <= u) lo = mid + 1; else hi = mid;
    }
    return lo < v ? lo : v - 1;
}

// one warp per document, doc_tokens <= 256: draw, warp-bitonic sort in smem, run-length encode
__global__ void __launch_bounds__(256) synth_sparse_kernel(uint64_t skey, int64_t row0, int64_t n, int vocab,
                                                           int doc_tokens, const uint64_t* __restrict__ thr,
                                                           const float* __restrict__ idf,
                                                           const float* __restrict__ tff, int64_t term_mul,
                                                           int64_t* __restrict__ counts,
                                                           const int64_t* __restrict__ indptr,
                                                           uint32_t* __restrict__ terms, float* __restrict__ wts) {
    __shared__ uint32_t tok[8][256];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t d = blockIdx.x * 8ll + wib;
    if (d >= n) return;
    const uint64_t rk = row_key(skey, (uint64_t)(row0 + d));
    uint32_t* t = tok[wib];
    for (int i = lane; i < 256; i += 32) {
        uint32_t v = 0xFFFFFFFFu;  // padding sorts last
        if (i < doc_tokens) {
            const int r = zipf_rank(thr, vocab, mix64(rk + (uint64_t)i) >> 11);
            v = (uint32_t)(((int64_t)r * term_mul) % vocab);
        }
        t[i] = v;
    }
    __syncwarp();
    // ascending bitonic sort of 256 u32
    for (int k = 2; k <= 256; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < 256; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = t[i], b = t[ixj];
                    const bool up = ((i & k) == 0);
                    if (up ? (a > b) : (a < b)) { t[i] = b; t[ixj] = a; }
                }
            }
            __syncwarp();
        }
    // lane owns positions [lane*8, lane*8+8): count run starts
    int starts = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = lane * 8 + k;
        if (i < doc_tokens && (i == 0 || t[i] != t[i - 1])) ++starts;
    }
    int incl = starts;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, dlt);
        if (lane >= dlt) incl += o;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (counts != nullptr) {
        if (lane == 0) counts[d] = total;
        return;
    }
    int64_t o = indptr[d] + (incl - starts);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = lane * 8 + k;
        if (i < doc_tokens && (i == 0 || t[i] != t[i - 1])) {
            int e = i + 1;
            while (e < doc_tokens && t[e] == t[i]) ++e;
            terms[o] = t[i];
            wts[o] = __fmul_rn(idf[t[i]], tff[e - i]);
            ++o;
        }
    }
}

int launch_synth_sparse(cudaStream_t st, uint64_t seed, int64_t row0, int64_t n, int vocab, int doc_tokens,
                        const uint64_t* thr, const float* idf, const float* tff, int64_t term_mul, int64_t* counts,
                        const int64_t* indptr, uint32_t* terms, float* w) {
    if (n <= 0) return B200RAG_OK;
    if (doc_tokens < 1 || doc_tokens > 256) { set_error("synth_sparse: doc_tokens must be in [1,256]"); return B200RAG_ERR_INVALID; }
    synth_sparse_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(stream_key(seed, 3), row0, n, vocab, doc_tokens, thr,
                                                                idf, tff, term_mul, counts, indptr, terms, w);
    B2_CUDA(cudaGetLastError());
    return B200RAG_OK;
}

int launch_exclusive_scan_i64(cudaStream_t st, const int64_t* in, int64_t n, int64_t* out) {
    // out[0..n] : out[0] = 0, out[i+1] = sum in[0..i]
    B2_CUDA(cudaMemsetAsync(out, 0, 8, st));
    if (n <= 0) return B200RAG_OK;
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, in, out + 1, n, st);
    void* tmp = nullptr;
    B2_CUDA(cudaMallocAsync(&tmp, tb + 16, st));
    cudaError_t e = cub::DeviceScan::InclusiveSum(tmp, tb, in, out + 1, n, st);
    cudaFreeAsync(tmp, st);
    if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceScan::InclusiveSum");
    return B200RAG_OK;
}

__global__ void synth_collection_mask_kernel(uint64_t skey, int64_t row0, int64_t n, const uint64_t* __restrict__ thr,
                                             int n_coll, int coll, uint32_t* __restrict__ out) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool bit = false;
    if (r < n) {
        const uint64_t u = mix64(row_key(skey, (uint64_t)(row0 + r))) >> 11;
        bit = zipf_rank(thr, n_coll, u) == coll;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31) == 0 && (r >> 5) < ((n + 31) >> 5)) out[r >> 5] = word;
}

int launch_synth_collection_mask(cudaStream_t st, uint64_t seed, int64_t row0, int64_t n, const uint64_t* thr,
                                 int n_coll, int coll, uint32_t* out_words) {
    if (n <= 0) return B200RAG_OK;
    const int64_t threads = ((n + 31) / 32) * 32;
    synth_collection_mask_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(stream_key(seed, 5), row0, n, thr,
                                                                                    n_coll, coll, out_words);
    B2_CUDA(cudaGetLastError());
    return B200RAG_OK;
}

}  // namespace b200rag
