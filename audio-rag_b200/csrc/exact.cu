// Exhaustive exact legs and row compaction: the two whole-shard passes that are NOT on the steady-state search path.
//
//   launch_exhaustive_leg   canonical score (the re-score kernels of select.cu) of EVERY eligible row,
//                           full radix sort of the exact keys, first L emitted under R5-R7.  It involves no approximate
//                           score, no candidate cut and therefore no slack guard: it is the always-exact fallback of
//                           b200rag_search when the guard never clears (massive ties), and an in-library cross-check of
//                           the scan kernels (b200rag_set_exhaustive).  Replaces, like the scans, the legs of
//                           client.query_points (src/audio_rag/retrieval/qdrant.py:281-332).
//   compact_rows            delete_collection (qdrant.py:354-363 drops the collection's storage): surviving rows of the
//                           dense matrix, the forward sparse index and the id map are gathered into fresh buffers in
//                           their old order; the inverted index is rebuilt from the forward index afterwards.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "common.cuh"
#include "engine.h"

namespace b200rag {

__global__ void fill_row_ids_kernel(int64_t* __restrict__ dst, int64_t first, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = first + i;
}

int launch_fill_row_ids(Shard* s, int64_t* dst, int64_t first_id, int64_t n) {
    if (n <= 0) return B200RAG_OK;
    fill_row_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(dst, first_id, n);
    B2_CUDA(cudaGetLastError());
    return B200RAG_OK;
}

// ---------------------------------------------------------------------------------------------- exhaustive leg
// key (score field 0) of every eligible row: the "candidate list" handed to the re-score kernels is the whole shard
__global__ void all_rows_keys_kernel(int64_t n, const uint32_t* __restrict__ mask, uint64_t* __restrict__ keys) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool ok = mask == nullptr || ((mask[i >> 5] >> (i & 31)) & 1u);
    keys[i] = ok ? make_key(0.f, (uint32_t)i) : 0ull;
}

__global__ void emit_sorted_kernel(const uint64_t* __restrict__ sorted, int64_t n, int L, int has_thr, float thr,
                                   const int64_t* __restrict__ row_ids, b200rag_cand* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    uint64_t k = i < n ? sorted[i] : 0ull;
    if (k != 0 && has_thr && key_score(k) < thr) k = 0;     // (keys are sorted: what the threshold drops is a suffix)
    b200rag_cand c;
    c.id = k != 0 ? row_ids[key_row(k)] : -1;
    c.score = k != 0 ? key_score(k) : 0.f;
    c.valid = k != 0 ? 1u : 0u;
    out[i] = c;
}

int launch_exhaustive_leg(Shard* s, bool sparse, int batch, int L, int has_thr, float thr, b200rag_cand* out) {
    const int64_t n = s->n_rows;
    cudaStream_t st = s->stream;
    B2_TRY(s->ws.ex_keys.ensure((size_t)n * 8, 0, st));
    B2_TRY(s->ws.ex_sorted.ensure((size_t)n * 8, 0, st));
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortKeysDescending(nullptr, temp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, n, 0, 64, st);
    B2_TRY(s->ws.ex_temp.ensure(temp_bytes + 16, 0, st));
    uint64_t* keys = s->ws.ex_keys.as<uint64_t>();
    uint64_t* exact = s->ws.ex_sorted.as<uint64_t>();
    for (int q = 0; q < batch; ++q) {
        const uint32_t* mask = s->h_masks.empty() ? nullptr : s->h_masks[(size_t)q];
        all_rows_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, mask, keys);
        B2_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
        if (sparse) B2_TRY(launch_rescore_sparse(s, 1, n, keys, exact, q, true));
        else B2_TRY(launch_rescore_dense(s, 1, n, keys, exact, q));
        size_t tb = temp_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortKeysDescending(s->ws.ex_temp.p, tb, exact, keys, n, 0, 64, st);
        if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortKeysDescending");
        emit_sorted_kernel<<<(L + 127) / 128, 128, 0, st>>>(keys, n, L, has_thr, thr, s->row_ids.as<int64_t>(),
                                                            out + (size_t)q * L);
        B2_CUDA(cudaGetLastError());
        s->stats.kernel_launches += 2;
    }
    return B200RAG_OK;
}

// ---------------------------------------------------------------------------------------------- compaction
__global__ void expand_keep_kernel(const uint32_t* __restrict__ words, int64_t n, int64_t* __restrict__ flag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    flag[i] = i < n ? (int64_t)((words[i >> 5] >> (i & 31)) & 1u) : 0;
}

// cnt[new index of row i] = postings of row i, for surviving rows; cnt[new_n] = 0
__global__ void kept_counts_kernel(const uint32_t* __restrict__ words, const int64_t* __restrict__ pos,
                                   const int64_t* __restrict__ fwd_ptr, int64_t n, int64_t new_n,
                                   int64_t* __restrict__ cnt) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i == n) { cnt[new_n] = 0; return; }
    if (i > n) return;
    if ((words[i >> 5] >> (i & 31)) & 1u) cnt[pos[i]] = fwd_ptr[i + 1] - fwd_ptr[i];
}

// one warp per old row: surviving rows move to their new position (dense row, postings, id)
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint32_t* __restrict__ words, const int64_t* __restrict__ pos,
                                                          int64_t n, int row_u4,
                                                          const uint4* __restrict__ dense, uint4* __restrict__ dense2,
                                                          const int64_t* __restrict__ fwd_ptr, const int64_t* __restrict__ ptr2,
                                                          const uint32_t* __restrict__ terms, uint32_t* __restrict__ terms2,
                                                          const float* __restrict__ w, float* __restrict__ w2,
                                                          const int64_t* __restrict__ ids, int64_t* __restrict__ ids2) {
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n || !((words[i >> 5] >> (i & 31)) & 1u)) return;
    const int64_t j = pos[i];
    const uint4* src = dense + (size_t)i * row_u4;
    uint4* dst = dense2 + (size_t)j * row_u4;
    for (int k = lane; k < row_u4; k += 32) dst[k] = src[k];
    const int64_t s0 = fwd_ptr[i], e0 = fwd_ptr[i + 1], d0 = ptr2[j];
    for (int64_t k = s0 + lane; k < e0; k += 32) { terms2[d0 + (k - s0)] = terms[k]; w2[d0 + (k - s0)] = w[k]; }
    if (lane == 0) ids2[j] = ids[i];
}

static int scan_i64(cudaStream_t st, const int64_t* in, int64_t* out, int64_t count, DevBuf& temp) {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, count, st);
    B2_TRY(temp.ensure(tb + 16, 0, st));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(temp.p, tb, in, out, count, st);
    if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceScan::ExclusiveSum");
    return B200RAG_OK;
}

int compact_rows(Shard* s, const uint32_t* keep) {
    cudaStream_t st = s->stream;
    const int64_t n = s->n_rows;
    DevBuf flag, pos, cnt, temp, dense2, ptr2, terms2, w2, ids2;
    int rc = B200RAG_OK;
    int64_t new_n = 0, new_nnz = 0;
    const unsigned g1 = (unsigned)((n + 1 + 255) / 256);
    auto fail = [&](int code) {
        flag.release(); pos.release(); cnt.release(); temp.release();
        dense2.release(); ptr2.release(); terms2.release(); w2.release(); ids2.release();
        return code;
    };
    if ((rc = flag.ensure((size_t)(n + 1) * 8, 0, st)) != B200RAG_OK) return fail(rc);
    if ((rc = pos.ensure((size_t)(n + 1) * 8, 0, st)) != B200RAG_OK) return fail(rc);
    expand_keep_kernel<<<g1, 256, 0, st>>>(keep, n, flag.as<int64_t>());
    if ((rc = scan_i64(st, flag.as<int64_t>(), pos.as<int64_t>(), n + 1, temp)) != B200RAG_OK) return fail(rc);
    {
        cudaError_t e = cudaMemcpyAsync(&new_n, pos.as<int64_t>() + n, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "compact: row count"));
    }
    if (new_n == n) return fail(B200RAG_OK);                    // nothing to drop
    // postings of the surviving rows -> new forward pointers
    if ((rc = cnt.ensure((size_t)(new_n + 1) * 8, 0, st)) != B200RAG_OK) return fail(rc);
    if ((rc = ptr2.ensure((size_t)(new_n + 1) * 8, 0, st)) != B200RAG_OK) return fail(rc);
    kept_counts_kernel<<<g1, 256, 0, st>>>(keep, pos.as<int64_t>(), s->fwd_ptr.as<int64_t>(), n, new_n, cnt.as<int64_t>());
    if ((rc = scan_i64(st, cnt.as<int64_t>(), ptr2.as<int64_t>(), new_n + 1, temp)) != B200RAG_OK) return fail(rc);
    {
        cudaError_t e = cudaMemcpyAsync(&new_nnz, ptr2.as<int64_t>() + new_n, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "compact: postings count"));
    }
    const size_t row_bytes = (size_t)s->dim * 2;
    if ((rc = dense2.ensure((size_t)std::max<int64_t>(new_n, 1) * row_bytes, 0, st)) != B200RAG_OK) return fail(rc);
    if ((rc = terms2.ensure((size_t)(new_nnz + 1) * 4, 0, st)) != B200RAG_OK) return fail(rc);
    if ((rc = w2.ensure((size_t)(new_nnz + 1) * 4, 0, st)) != B200RAG_OK) return fail(rc);
    if ((rc = ids2.ensure((size_t)std::max<int64_t>(new_n, 1) * 8, 0, st)) != B200RAG_OK) return fail(rc);
    gather_rows_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(
        keep, pos.as<int64_t>(), n, (int)(row_bytes / 16), s->dense.as<uint4>(), dense2.as<uint4>(),
        s->fwd_ptr.as<int64_t>(), ptr2.as<int64_t>(), s->fwd_terms.as<uint32_t>(), terms2.as<uint32_t>(),
        s->fwd_w.as<float>(), w2.as<float>(), s->row_ids.as<int64_t>(), ids2.as<int64_t>());
    {
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "compact: gather"));
    }
    std::swap(s->dense, dense2);
    std::swap(s->fwd_ptr, ptr2);
    std::swap(s->fwd_terms, terms2);
    std::swap(s->fwd_w, w2);
    std::swap(s->row_ids, ids2);
    s->n_rows = new_n;
    s->nnz = new_nnz;
    s->built_rows = 0; s->n_blocks = 0; s->inv_nnz = 0;
    s->h_blk_base.clear();
    s->w_absmax = 0.f; s->wmax_nnz = 0;
    return fail(B200RAG_OK);                                    // releases the old buffers and the scratch
}

}  // namespace b200rag
