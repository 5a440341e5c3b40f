// Shared device helpers: ordered 64-bit candidate keys, bitonic selection networks, mbarrier / bulk-copy PTX.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200rag {

// ----------------------------------------------------------------------------------------------------
// Candidate key.  One u64 whose unsigned order is the leg order of SURVEY R5:
//   score descending, ties -> smaller row id first.
// high 32 bits: order-preserving transform of the fp32 score, low 32 bits: ~local_row.
// 0 is "no candidate": every real key is > 0 because ord(-inf) = 0x007FFFFF.
// ----------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t ord_f32(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float unord_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
    return ((uint64_t)ord_f32(score) << 32) | (uint64_t)(~row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~(uint32_t)k; }
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return unord_f32((uint32_t)(k >> 32)); }

__host__ __device__ __forceinline__ int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------------
// Bitonic sort, descending, of n (power of two) u64 keys in shared memory.
// warp flavour: one warp, __syncwarp between stages.  cta flavour: `nthreads` threads that all call it,
// synchronised on named barrier `bar_id` (0 == __syncthreads-compatible barrier for the whole CTA).
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_step(uint64_t* keys, int i, int j, int k) {
    int ixj = i ^ j;
    if (ixj > i) {
        uint64_t a = keys[i], b = keys[ixj];
        bool desc = ((i & k) == 0);
        if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
    }
}

__device__ __forceinline__ void warp_bitonic_desc(uint64_t* keys, int n, int lane) {
    __syncwarp();
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += 32) bitonic_step(keys, i, j, k);
            __syncwarp();
        }
}

__device__ __forceinline__ void named_bar_sync(int bar_id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void cta_bitonic_desc(uint64_t* keys, int n, int tid, int nthreads, int bar_id) {
    // Compare distances j < 64 stay inside aligned 64-key chunks: a warp owns whole chunks and only needs __syncwarp
    // there.  Only the j >= 64 steps cross warps and take a block barrier: 14 barriers instead of 55 for n = 1024.
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    named_bar_sync(bar_id, nthreads);
    if (n < 64 || nwarps == 0) {
        for (int k = 2; k <= n; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < n; i += nthreads) bitonic_step(keys, i, j, k);
                named_bar_sync(bar_id, nthreads);
            }
        return;
    }
    const int nchunks = n >> 6;
    for (int k = 2; k <= n; k <<= 1) {
        int j = k >> 1;
        if (j >= 64) {
            named_bar_sync(bar_id, nthreads);          // warp-local results of the previous phase become visible
            for (; j >= 64; j >>= 1) {
                for (int i = tid; i < n; i += nthreads) bitonic_step(keys, i, j, k);
                named_bar_sync(bar_id, nthreads);
            }
        }
        for (; j > 0; j >>= 1) {
            for (int c = warp; c < nchunks; c += nwarps) {
                bitonic_step(keys, (c << 6) + lane, j, k);
                bitonic_step(keys, (c << 6) + 32 + lane, j, k);
            }
            __syncwarp();
        }
    }
    named_bar_sync(bar_id, nthreads);
}

// ----------------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA engine, 1-D form).  SASS: SYNCS.* / UBLKCP.
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; bytes multiple of 16, both addresses 16-byte aligned; completes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
#endif  // __CUDACC__

}  // namespace b200rag
