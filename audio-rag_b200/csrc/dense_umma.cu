// K1b  dense_gemm: query-batch x corpus inner products on the 5th-gen tensor cores (tcgen05 + TMEM), corpus and
// query tiles staged by TMA, top-k fused in the TMEM epilogue -- the score matrix never reaches HBM.
//
// Replaces the dense leg of client.query_points for BATCHES of queries (search_batch; reference call sites
// src/audio_rag/retrieval/qdrant.py:285-288, 317-332 issue one query at a time).
//
// GEMM shape per CTA tile:  D[128 queries x 256 corpus rows] (fp32, TMEM) = Q[128 x K] * C[256 x K]^T,
// both operands bf16 K-major in shared memory with the 128-byte swizzle, K = dim stepped 64 elements per
// pipeline stage (16 KB of queries + 32 KB of corpus rows per stage, 4 stages), 4 x tcgen05.mma (K = 16) per stage.
// With more than 128 queries in a pass the kernel runs as CTA PAIRS (cluster of 2, tcgen05.mma.cta_group::2):
// M = 256 queries, each CTA stages its own 128 queries and half of the corpus tile (32 KB per stage, 6 stages).
// Queries are the M dimension on purpose: TMEM lane i then holds query i, so ONE epilogue thread owns ONE query and
// keeps that query's running threshold in a register: the common case per score is one FSETP.  Three epilogues:
// a sorted candidate list in registers (top-k up to 64), replace-min lists in shared memory (retry path), and a
// stateless FILTER against a fixed threshold from a 1/16 sample pass (any top-k, 128-256 queries per pass).
// Thresholds are shared grid-wide through a monotone atomicMax like in the SIMT scan.
//
// Warp roles (256 threads, 1 CTA/SM, persistent over 256-row corpus tiles):
//   warp 0  TMA producer   (one lane): cp.async.bulk.tensor.2d of the query k-block and the corpus k-block
//   warp 1  MMA issuer     (one lane, leader CTA of a pair): tcgen05.mma, tcgen05.commit -> smem-empty / tmem-full
//   warp 2  TMEM allocator (512 columns = two 256-column accumulator buffers, so the epilogue of tile i overlaps
//                           the MMAs of tile i+1)
//   warps 4-7 epilogue: tcgen05.ld 32 lanes x 32 columns at a time, compare, insert / append
//
// Roofline: HBM for <= ~200 queries per pass (intensity = B flop/byte), tensor pipe above.
// Algorithmic bytes per launch = n_rows * dim * 2.
#include <cuda.h>

#include "common.cuh"
#include "engine.h"

namespace b200rag {

constexpr int kGemmM = 128;        // queries per pass (TMEM lanes)
constexpr int kGemmN = 256;        // corpus rows per tile (TMEM columns per accumulator buffer)
constexpr int kGemmKB = 64;        // K elements per pipeline stage (128 bytes: one swizzle span)
constexpr int kGemmStages = 4;
constexpr int kGemmThreads = 256;
constexpr uint32_t kStageABytes = kGemmM * kGemmKB * 2;   // 16 KB
constexpr uint32_t kStageBBytes = kGemmN * kGemmKB * 2;   // 32 KB
constexpr uint32_t kStageBytes = kStageABytes + kStageBBytes;
constexpr uint32_t kTmemCols = 512;
// CTA pair (cta_group::2): M = 256 queries (128 per CTA), N = 256 corpus rows split 128 + 128 between the two CTAs
constexpr uint32_t kStageBBytesPair = (kGemmN / 2) * kGemmKB * 2;   // 16 KB
constexpr uint32_t kStageBytesPair = kStageABytes + kStageBBytesPair;
constexpr int kGemmStagesPair = 7;

struct GemmParams {
    const void* corpus;            // for the contiguous L2 prefetch
    int64_t n_rows;
    int64_t n_tiles;
    int dim;
    int batch;                     // valid queries in this pass (<= 128)
    const uint32_t* const* masks;  // device array [batch] of eligibility bitmaps (entries may be null) or nullptr
    uint64_t* g_thr;               // [batch]
    uint64_t* out;                 // [batch][grid][Lc]
    int64_t out_q_stride;          // grid * Lc
    int Lc;
    int stages;                    // pipeline depth actually used (2..kGemmStages), set by the launcher
    int tile_stride;               // 1 = every 256-row tile; 16 = the sample pass (tiles 0, 16, 32, ...)
    uint64_t* pool;                // filter mode: [batch][pool_cap] keys at or above the query's fixed threshold
    int* pool_cnt;                 // filter mode: [batch] append counters (may exceed pool_cap: overflow is detected later)
    int pool_cap;
    float* dbg_scores;             // optional [batch][n_rows] raw approximate scores (tests)
    int l2_prefetch;               // prefetch every corpus tile into L2 one tile ahead of its loads
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// corpus tile -> L2 only (no shared-memory destination, no barrier): issued one tile ahead of the loads, so that the
// loads that fill the pipeline stages find their bytes in L2 (~1 us) instead of HBM (~2 us under load); with 6 stages a
// freed stage has ~2 us of MMAs ahead of it, and the MMA thread spent 15 % of its time waiting for `full` barriers
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// ---- CTA-pair (cta_group::2) variants.  The pair's leader is the CTA whose shared-window address has the peer bit
// (bit 24) clear; masking an mbarrier address with kPeerMask names the LEADER's copy of that barrier from either CTA.
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once the pair's MMAs so far have retired) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
// arrive on the LEADER's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// v[j] for a run-time j without local memory: a 5-level select tree over the 32 registers (31 selects)
__device__ __forceinline__ float pick32(const uint32_t (&v)[32], int j) {
    uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
    return __uint_as_float((j & 16) ? d[1] : d[0]) + 0.0f;
}
// one lane of a fully active warp (elect.sync): the warp runs the loop, the elected lane issues the TMA / MMA instructions.
// With the loop warp-uniform the compiler keeps descriptors, barrier addresses and coordinates in UNIFORM registers; inside
// an `if (lane == 0)` region it wrapped every UTCHMMA in an ELECT + 5 x R2UR.BROADCAST "waterfall" loop (~15 dependent
// instructions per MMA: the issue loop then took about as long per k-block as the tensor pipe needs for its 4 MMAs).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Candidate lists in shared memory, one per query: an UNSORTED set of the Lc best keys seen so far plus, in the
// owning lane's registers, the fill count and the position/value of the current minimum (= the running threshold).
// An insert overwrites the minimum and rescans for the new one: Lp/2 independent 128-bit loads, no shifting, so
// the cost does not depend on where the key ranks and up to 32 lanes insert in the same round.  Lists are sorted
// once, at the end of the kernel.  Lp = Lc rounded up to even (a pad slot holds ~0 and is never the minimum).
__device__ __forceinline__ void lane_replace(uint64_t* lst, int Lc, int Lp, uint64_t key, int& cnt, int& minpos,
                                             uint64_t& lmin) {
    if (cnt < Lc) {
        lst[cnt++] = key;
        if (cnt < Lc) return;
    } else {
        lst[minpos] = key;
    }
    uint64_t m = ~0ull;
    int mp = 0;
    const ulonglong2* l2 = reinterpret_cast<const ulonglong2*>(lst);
#pragma unroll 4
    for (int i = 0; i < Lp / 2; ++i) {
        const ulonglong2 kk = l2[i];
        if (kk.x < m) { m = kk.x; mp = 2 * i; }
        if (kk.y < m) { m = kk.y; mp = 2 * i + 1; }
    }
    minpos = mp;
    lmin = m;
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row atoms 1024 bytes apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGemmN >> 3) << 17) |
                            ((uint32_t)(kGemmM >> 4) << 24);
constexpr uint32_t kIdescPair = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGemmN >> 3) << 17) |
                                ((uint32_t)((2 * kGemmM) >> 4) << 24);   // M = 256 across the pair

// LREG > 0: each epilogue lane keeps its query's candidate list SORTED IN REGISTERS (LREG keys; Lc <= LREG; the
//           threshold is the LREG-th best, slightly weaker than the Lc-th, which only admits a few more inserts).
//           A sorted insert is LREG independent compare/selects: no memory, no dependent chain.
// LREG == 0: lists live in shared memory (unsorted, replace-min) for large top-k.
// LREG < 0:  FILTER mode, no lists at all: g_thr[q] holds a FIXED threshold (the K_s-th best score of a 1/16 sample of
//            the shard, see launch_dense_gemm_filtered); every score at or above it is appended to the query's pool in
//            global memory (~16 K_s keys per query over the whole shard).  The epilogue carries no per-query state
//            besides one register, so 128 queries per pass fit whatever the top-k (smem lists: 32 at top-100).
// CG == 2: the kernel runs as CTA PAIRS (cluster of 2, cta_group::2): one tcgen05.mma covers 256 queries x 256 rows,
//          each CTA stages its own 128 queries and HALF of the corpus tile, so a corpus row crosses L2 -> SM once
//          per 256 queries instead of once per 128 and a whole batch of 256 needs ONE pass over HBM.  The leader CTA
//          issues the MMAs; commits are multicast to both CTAs' barriers; each CTA's epilogue reads its own TMEM.
template <int LREG, int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
dense_gemm_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c,
                  const GemmParams p) {
    constexpr uint32_t STAGE_BYTES = CG == 2 ? kStageBytesPair : kStageBytes;
    constexpr int MAX_STAGES = CG == 2 ? kGemmStagesPair : kGemmStages;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    extern __shared__ uint8_t gsm_raw[];
    // 1024-byte alignment for the 128B-swizzled tiles
    // (offset arithmetic on the shared window keeps the pointer in the shared address space: LDS/STS, not generic)
    uint8_t* gsm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
    const int S = p.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(gsm + (size_t)S * STAGE_BYTES);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tfull = empty + MAX_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tempty + 2);   // (16 u64 barrier slots precede: 16-byte aligned)
    uint64_t* lists_s = reinterpret_cast<uint64_t*>(tmem_base_s + 4);             // [ceil32(batch)][Lp], 16-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.dim / kGemmKB;
    const int64_t unit = (int64_t)blockIdx.x / CG, units = (int64_t)gridDim.x / CG;   // CTA (or CTA pair) and their number
    const int64_t t0 = p.n_tiles * unit / units;
    const int64_t t1 = p.n_tiles * (unit + 1) / units;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_q);
        prefetch_tmap(&map_c);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4 * CG); }
        fence_mbar_init();
    }
    if (warp == 2) {
        if (CG == 2) tmem_alloc_pair(tmem_base_s, kTmemCols);
        else tmem_alloc(tmem_base_s, kTmemCols);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------ TMA producer
        // (the whole warp runs the loop, one elected lane issues: see elect_one)
        const bool elected = elect_one();
        int st = 0;
        uint32_t ph = 0;
        for (int64_t t = t0; t < t1; ++t) {
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty[st], ph ^ 1u);
                uint8_t* sa = gsm + (size_t)st * STAGE_BYTES;
                if (elected) {
                    const int row0 = (int)(t * p.tile_stride * kGemmN) + (CG == 2 ? (int)rank * (kGemmN / 2) : 0);
                    if (p.l2_prefetch && t + 1 < t1)
                        tma_prefetch_2d(&map_c, kb * kGemmKB, row0 + p.tile_stride * kGemmN);     // the same k-block of my next tile
                    if (CG == 2) {
                        // both CTAs' copies complete on the LEADER's full barrier (which expects the pair's bytes)
                        if (leader) mbar_arrive_expect_tx(&full[st], 2 * STAGE_BYTES);
                        tma_load_2d_pair(sa, &map_q, &full[st], kb * kGemmKB, (int)rank * kGemmM);
                        tma_load_2d_pair(sa + kStageABytes, &map_c, &full[st], kb * kGemmKB, row0);
                    } else {
                        mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
                        tma_load_2d(sa, &map_q, &full[st], kb * kGemmKB, 0);
                        tma_load_2d(sa + kStageABytes, &map_c, &full[st], kb * kGemmKB, row0);
                    }
                }
                __syncwarp();
                if (++st == S) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------ MMA issuer
        if (leader) {
            const bool elected = elect_one();
            int st = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int64_t t = t0; t < t1; ++t, ++it) {
                const int buf = it & 1;
                mbar_wait(&tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kGemmN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(gsm + (size_t)st * STAGE_BYTES);
                    const uint64_t da = make_sw128_desc(sa);
                    const uint64_t db = make_sw128_desc(sa + kStageABytes);
                    if (elected) {
#pragma unroll
                        for (int k = 0; k < kGemmKB / 16; ++k) {
                            if (CG == 2) umma_bf16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdescPair,
                                                        (kb | k) != 0 ? 1u : 0u);
                            else umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc,
                                           (kb | k) != 0 ? 1u : 0u);
                        }
                        // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                        if (CG == 2) umma_commit_pair(&empty[st]); else umma_commit(&empty[st]);
                        if (kb == nkb - 1) { if (CG == 2) umma_commit_pair(&tfull[buf]); else umma_commit(&tfull[buf]); }
                    }
                    __syncwarp();
                    if (++st == S) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------------------------ epilogue
        const int quad = warp - 4;                       // == warp % 4: the TMEM lane quadrant this warp may read
        const int qi = (int)rank * kGemmM + quad * 32 + lane;   // my query
        const bool active = qi < p.batch;
        const bool warp_active = (int)rank * kGemmM + quad * 32 < p.batch;
        const int Lp = (p.Lc + 1) & ~1;
        uint64_t* wlists = lists_s + (size_t)(quad * 32) * Lp;     // this warp's 32 lists (LREG == 0 only)
        const uint32_t* mask = (active && p.masks != nullptr) ? p.masks[qi] : nullptr;
        uint64_t L[LREG > 0 ? LREG : 1];
#pragma unroll
        for (int i = 0; i < (LREG > 0 ? LREG : 1); ++i) L[i] = 0;
        if (LREG == 0 && warp_active) {
            for (int i = lane; i < 32 * Lp; i += 32) wlists[i] = 0;
            __syncwarp();
            if (Lp != p.Lc) wlists[(size_t)lane * Lp + p.Lc] = ~0ull;
        }
        __syncwarp();
        uint64_t thr = 0;          // acceptance threshold: max(my list's minimum once full, grid-wide threshold)
        uint64_t lmin = 0;         // my list's minimum (0 until the list is full)
        int cnt = 0, minpos = 0;
        float thr_s = -INFINITY;
        int it = 0;
        for (int64_t t = t0; t < t1; ++t, ++it) {
            const int buf = it & 1;
            if (active) {
                const uint64_t g = ld_volatile_u64(&p.g_thr[qi]);
                if (g > thr) { thr = g; thr_s = key_score(g); }
            }
            // eligibility of the tile's 256 rows for my query = 8 consecutive words: pull their line towards L1 now,
            // read one word per chunk below
            const int64_t tt = t * p.tile_stride;          // the corpus tile this iteration scores
            if (mask != nullptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(mask + tt * (kGemmN / 32)));
            mbar_wait(&tfull[buf], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kGemmN);
            for (int c = 0; c < kGemmN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)(c * 32), v);
                tmem_ld_wait();
                if (!warp_active) continue;
                const int64_t row0 = tt * kGemmN + c * 32;
                if (p.dbg_scores != nullptr && active) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (row0 + j < p.n_rows)
                            p.dbg_scores[(size_t)qi * p.n_rows + row0 + j] = __uint_as_float(v[j]) + 0.0f;
                }
                // phase A: one compare per score against my query's threshold (registers only)
                uint32_t cm = 0;
                if (active) {
                    // one compare + one bit-insert per score; rows beyond the shard's end (last tile only) and ineligible
                    // rows are cut from the whole 32-bit word afterwards
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        cm |= (__uint_as_float(v[j]) >= thr_s) ? (1u << j) : 0u;
                    const int64_t left = p.n_rows - row0;
                    if (left < 32) cm &= left > 0 ? ((1u << (int)left) - 1u) : 0u;
                    if (mask != nullptr) cm &= __ldg(mask + tt * (kGemmN / 32) + c);
                }
                if (!__any_sync(0xffffffffu, cm != 0)) continue;
                // phase B (rare after warm-up): insert survivors.  A survivor's score is picked out of the chunk's 32
                // registers by a select tree (pick32): round 1 staged every chunk's scores in 16.9 KB of shared memory for
                // this, which is now a 7th pipeline stage for the CTA pairs.
                // every lane with survivors inserts one of its own per round (up to 32 inserts per round)
                while (__any_sync(0xffffffffu, cm != 0)) {
                    if (cm != 0) {
                        const int j = __ffs(cm) - 1;
                        cm &= cm - 1;
                        const int64_t row = row0 + j;
                        const uint64_t key = make_key(pick32(v, j), (uint32_t)row);
                        const bool ok = LREG < 0 ? key >= thr : key > thr;   // (eligibility was applied to cm already)
                        if constexpr (LREG < 0) {
                            if (ok) {
                                const int pos = atomicAdd(&p.pool_cnt[qi], 1);
                                if (pos < p.pool_cap) p.pool[(size_t)qi * p.pool_cap + pos] = key;
                            }
                        } else if (ok) {
                            if constexpr (LREG > 0) {
                                // sorted insert (descending), every element decided from OLD neighbours: no chain
#pragma unroll
                                for (int i = LREG - 1; i >= 1; --i) {
                                    const bool gi = key > L[i], gm = key > L[i - 1];
                                    L[i] = gi ? (gm ? L[i - 1] : key) : L[i];
                                }
                                L[0] = key > L[0] ? key : L[0];
                                lmin = L[LREG - 1];
                            } else {
                                lane_replace(wlists + (size_t)lane * Lp, p.Lc, Lp, key, cnt, minpos, lmin);
                            }
                            if (lmin > thr) {
                                thr = lmin;
                                thr_s = key_score(lmin);
                                atomicMax(reinterpret_cast<unsigned long long*>(&p.g_thr[qi]), (unsigned long long)lmin);
                            }
                        }
                        // survivors flagged against the chunk-start threshold that the new threshold rules out
                        while (cm != 0 && pick32(v, __ffs(cm) - 1) < thr_s) cm &= cm - 1;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_leader(&tempty[buf]); else mbar_arrive(&tempty[buf]); }
        }
        __syncwarp();
        if constexpr (LREG > 0) {
            // my list is already sorted: first Lc keys -> global [q][cta][Lc]
            if (active) {
                uint64_t* o = p.out + (size_t)qi * p.out_q_stride + (size_t)unit * p.Lc;   // one list per CTA (pair)
#pragma unroll
                for (int i = 0; i < LREG; ++i)
                    if (i < p.Lc) o[i] = L[i];
            }
        } else if (LREG == 0 && warp_active) {
            // unsorted lists -> global [q][cta][Lc]; the merge tree sorts (it never assumes order)
            for (int ql = 0; ql < 32; ++ql) {
                const int q = quad * 32 + ql;
                if (q >= p.batch) break;
                const uint64_t* src_l = wlists + (size_t)ql * Lp;
                uint64_t* o = p.out + (size_t)q * p.out_q_stride + (size_t)blockIdx.x * p.Lc;
                for (int i = lane; i < p.Lc; i += 32) o[i] = src_l[i];
            }
        }
    }

    // ---- teardown: everyone meets, then the allocating warp frees TMEM
    tc_fence_before();
    if (CG == 2) cluster_sync_all();   // the peer may still be reading operands / signalling this CTA's barriers
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static int make_map(CUtensorMap* map, const void* base, int64_t rows, int dim, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B200RAG_ERR_CUDA; }
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
    cuuint32_t box[2] = {(cuuint32_t)kGemmKB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r)); return B200RAG_ERR_CUDA; }
    return B200RAG_OK;
}

int dense_gemm_nlists(const Shard* s) { return s->sm_count; }

template <int LREG, int CG>
static int launch_gemm_t(Shard* s, const CUtensorMap& map_q, const CUtensorMap& map_c, const GemmParams& p, int grid,
                         size_t smem) {
    auto kern = dense_gemm_kernel<LREG, CG>;
    static AttrCache attr;
    if (attr.raise(s->cfg.device, 227 * 1024))
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = CG == 2 ? 1 : 0;
    B2_CUDA(cudaLaunchKernelEx(&cfg, kern, map_q, map_c, p));
    return B200RAG_OK;
}

// One or more corpus passes.  A pass takes up to 256 queries on CTA pairs (cta_group::2) when more than 128 remain
// and the epilogue keeps no shared-memory lists, else up to 128 on single CTAs.  tile_stride > 1 scores only every
// tile_stride-th 256-row tile (the sample pass); pool != nullptr selects the filter epilogue (out_lists unused).
// List layout: out_lists[q][nl][Lc] with nl = *nlists for EVERY query (passes on pairs fill only nl/2... see below).
static int gemm_passes(Shard* s, int batch, int Lc, uint64_t* out_lists, int* nlists, float* dbg_scores, int tile_stride,
                       uint64_t* pool, int* pool_cnt, int pool_cap) {
    const int64_t all_tiles = (s->n_rows + kGemmN - 1) / kGemmN;
    const int64_t n_tiles = (all_tiles + tile_stride - 1) / tile_stride;
    int grid1 = s->sm_count;                       // single-CTA passes
    if (n_tiles < grid1) grid1 = (int)(n_tiles > 0 ? n_tiles : 1);
    int units2 = s->sm_count / 2;                  // CTA-pair passes
    if (n_tiles < units2) units2 = (int)(n_tiles > 0 ? n_tiles : 1);
    if (nlists != nullptr) *nlists = grid1;        // slots per query; a pair pass writes units2 <= grid1 of them, zeros the rest
    const bool filter = pool != nullptr;
    const int lreg = filter ? -1 : (Lc <= 32 ? 32 : (Lc <= 64 ? 64 : 0));
    const bool pairs_ok = s->gemm_pairs && lreg != 0 && dbg_scores == nullptr;
    // shared memory: S pipeline stages + score staging (+ one list of Lc keys per query when the lists do not fit
    // in registers)
    const size_t max_smem = 227 * 1024, fixed = 1024 + 512;
    const int Lp = (Lc + 1) & ~1;
    const size_t per_q = lreg != 0 ? 0 : (size_t)Lp * 8;
    int qpp = kGemmM;                                   // queries per single-CTA pass
    while (qpp > 32 && fixed + 2 * (size_t)kStageBytes + (size_t)qpp * per_q > max_smem) qpp -= 32;
    if (fixed + 2 * (size_t)kStageBytes + (size_t)qpp * per_q > max_smem) {
        set_error("dense_gemm: top-k too large for shared memory");
        return B200RAG_ERR_INVALID;
    }
    CUtensorMap map_c, map_c2;
    B2_TRY(make_map(&map_c, s->dense.p, s->n_rows, s->dim, kGemmN));
    if (pairs_ok) B2_TRY(make_map(&map_c2, s->dense.p, s->n_rows, s->dim, kGemmN / 2));
    int q0 = 0;
    while (q0 < batch) {
        const bool pair = pairs_ok && batch - q0 > kGemmM;
        const int cap_q = pair ? 2 * kGemmM : qpp;
        const int nq = batch - q0 < cap_q ? batch - q0 : cap_q;
        const size_t list_bytes = pair ? 0 : (size_t)((nq + 31) / 32 * 32) * per_q;
        const size_t stage_bytes = pair ? kStageBytesPair : kStageBytes;
        int stages = (int)((max_smem - fixed - list_bytes) / stage_bytes);
        const int max_stages = pair ? kGemmStagesPair : kGemmStages;
        if (stages > max_stages) stages = max_stages;
        if (s->gemm_stage_cap > 1 && stages > s->gemm_stage_cap) stages = s->gemm_stage_cap;
        // a hybrid batch whose sparse leg runs concurrently: leave ~64 KB of the SM's shared memory to its CTAs
        if (s->gemm_smem_reserve > 0)
            while (stages > 2 && fixed + (size_t)stages * stage_bytes + list_bytes + s->gemm_smem_reserve > max_smem) --stages;
        const size_t smem = fixed + (size_t)stages * stage_bytes + list_bytes;
        const int grid = pair ? 2 * units2 : grid1;
        CUtensorMap map_q;
        B2_TRY(make_map(&map_q, s->ws.q_bits.as<uint16_t>() + (size_t)q0 * s->dim, nq, s->dim, kGemmM));
        GemmParams p{};
        p.corpus = s->dense.p;
        p.n_rows = s->n_rows;
        p.n_tiles = n_tiles;
        p.dim = s->dim;
        p.batch = nq;
        p.masks = s->h_masks.empty() ? nullptr : s->ws.q_masks.as<const uint32_t*>() + q0;
        p.g_thr = s->ws.thr.as<uint64_t>() + q0;
        p.out = filter ? nullptr : out_lists + (size_t)q0 * grid1 * Lc;
        p.out_q_stride = (int64_t)grid1 * Lc;
        p.Lc = Lc;
        p.stages = stages;
        p.tile_stride = tile_stride;
        p.pool = filter ? pool + (size_t)q0 * pool_cap : nullptr;
        p.pool_cnt = filter ? pool_cnt + q0 : nullptr;
        p.pool_cap = pool_cap;
        p.dbg_scores = dbg_scores != nullptr ? dbg_scores + (size_t)q0 * s->n_rows : nullptr;
        p.l2_prefetch = s->gemm_l2_prefetch ? 1 : 0;
        if (pair && !filter && units2 < grid1) {
            // a pair pass fills list slots [0, units2) of each of its queries: the others must read as "no candidate"
            B2_CUDA(cudaMemsetAsync(p.out, 0, (size_t)nq * grid1 * Lc * 8, s->stream));
        }
        if (pair) {
            if (lreg < 0) B2_TRY((launch_gemm_t<-1, 2>(s, map_q, map_c2, p, grid, smem)));
            else if (lreg == 32) B2_TRY((launch_gemm_t<32, 2>(s, map_q, map_c2, p, grid, smem)));
            else B2_TRY((launch_gemm_t<64, 2>(s, map_q, map_c2, p, grid, smem)));
        } else {
            if (lreg < 0) B2_TRY((launch_gemm_t<-1, 1>(s, map_q, map_c, p, grid, smem)));
            else if (lreg == 32) B2_TRY((launch_gemm_t<32, 1>(s, map_q, map_c, p, grid, smem)));
            else if (lreg == 64) B2_TRY((launch_gemm_t<64, 1>(s, map_q, map_c, p, grid, smem)));
            else B2_TRY((launch_gemm_t<0, 1>(s, map_q, map_c, p, grid, smem)));
        }
        s->stats.kernel_launches++;
        if (tile_stride == 1) s->stats.dense_passes++;
        q0 += nq;
    }
    return B200RAG_OK;
}

int launch_dense_gemm(Shard* s, int batch, int Lc, uint64_t* out_lists, int* nlists, float* dbg_scores) {
    s->stats.dense_path = 2;
    s->stats.dense_passes = 0;
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[0], s->stream)); }
    B2_TRY(gemm_passes(s, batch, Lc, out_lists, nlists, dbg_scores, 1, nullptr, nullptr, 0));
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[1], s->stream)); s->ev_dense = true; }
    s->stats.dense_bytes = (int64_t)s->stats.dense_passes * s->n_rows * s->dim * 2;
    return B200RAG_OK;
}

// ---- large top-k with query batches: sample pass -> fixed thresholds -> filter pass -> pool selection -------------
constexpr int kSampleStride = 16;

__global__ void set_filter_thr_kernel(const uint64_t* __restrict__ merged, int Ks, uint64_t* __restrict__ g_thr,
                                      int* __restrict__ pool_cnt, int batch) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= batch) return;
    g_thr[q] = merged[(size_t)q * Ks + Ks - 1];   // 0 while the sample holds fewer than Ks eligible rows: accept everything
    pool_cnt[q] = 0;
}

// one CTA per query: sort the pool, emit the best Lc keys; flag the search as ambiguous (-> robust retry) when the
// fixed threshold turned out too high (fewer than Lc rows passed although a threshold was applied) or the pool overflowed
__global__ void __launch_bounds__(512) pool_select_kernel(const uint64_t* __restrict__ pool,
                                                          const int* __restrict__ pool_cnt, int cap,
                                                          const uint64_t* __restrict__ g_thr, int Lc,
                                                          uint64_t* __restrict__ approx, int32_t* __restrict__ ambiguous) {
    extern __shared__ __align__(16) uint64_t pkeys[];
    const int q = blockIdx.x;
    const int cnt = pool_cnt[q];
    const int n = cnt < cap ? cnt : cap;
    int npow2 = next_pow2(n > Lc ? n : Lc);
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) pkeys[i] = i < n ? pool[(size_t)q * cap + i] : 0ull;
    cta_bitonic_desc(pkeys, npow2, threadIdx.x, blockDim.x, 0);
    for (int i = threadIdx.x; i < Lc; i += blockDim.x) approx[(size_t)q * Lc + i] = pkeys[i];
    if (threadIdx.x == 0 && ambiguous != nullptr && (cnt > cap || (cnt < Lc && g_thr[q] != 0))) atomicAdd(ambiguous, 1);
}

int dense_filter_sample_k(int Lc) {
    int ks = (5 * Lc + 31) / 32;     // ~2.5 Lc / 16: the expected pool is 2.5 Lc rows, > 4 sigma above Lc
    if (ks < 16) ks = 16;
    if (ks > 64) ks = 64;
    return ks;
}

int launch_dense_gemm_filtered(Shard* s, int batch, int Lc, uint64_t* scratch_a, uint64_t* scratch_b, uint64_t* approx,
                               int32_t* ambiguous) {
    const int Ks = dense_filter_sample_k(Lc);
    int cap = next_pow2(4 * kSampleStride * Ks);
    if (cap < 1024) cap = 1024;
    if (cap > 8192) cap = 8192;
    B2_TRY(s->ws.pool.ensure((size_t)batch * cap * 8 + (size_t)batch * 4, 0, s->stream));
    uint64_t* pool = s->ws.pool.as<uint64_t>();
    int* pool_cnt = reinterpret_cast<int*>(pool + (size_t)batch * cap);
    s->stats.dense_path = 2;
    s->stats.dense_passes = 0;
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[0], s->stream)); }
    // 1. sample pass: top-Ks per query over every 16th tile (register lists, dynamic thresholds in g_thr)
    int nl = 0;
    B2_TRY(gemm_passes(s, batch, Ks, scratch_a, &nl, nullptr, kSampleStride, nullptr, nullptr, 0));
    uint64_t* merged = nullptr;
    B2_TRY(launch_merge_tree(s, batch, nl, Ks, scratch_a, scratch_b, &merged));
    // 2. fixed threshold per query = Ks-th best of the sample
    set_filter_thr_kernel<<<(batch + 127) / 128, 128, 0, s->stream>>>(merged, Ks, s->ws.thr.as<uint64_t>(), pool_cnt, batch);
    B2_CUDA(cudaGetLastError());
    // 3. filter pass over the whole shard
    B2_TRY(gemm_passes(s, batch, Lc, nullptr, nullptr, nullptr, 1, pool, pool_cnt, cap));
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[1], s->stream)); s->ev_dense = true; }
    // 4. best Lc of each pool
    static AttrCache attr;
    if (attr.raise(s->cfg.device, 8192 * 8))
        B2_CUDA(cudaFuncSetAttribute(pool_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
    const int npow2_max = next_pow2(cap > Lc ? cap : Lc);
    pool_select_kernel<<<batch, 512, (size_t)npow2_max * 8, s->stream>>>(pool, pool_cnt, cap, s->ws.thr.as<uint64_t>(), Lc,
                                                                        approx, ambiguous);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 2;
    s->stats.dense_bytes = (int64_t)s->stats.dense_passes * s->n_rows * s->dim * 2;
    return B200RAG_OK;
}

}  // namespace b200rag
