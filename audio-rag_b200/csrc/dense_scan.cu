// K1a  dense_scan: bandwidth-bound exact inner-product scan for small query groups (1-2 queries per pass).
//
// Replaces the dense leg of client.query_points (reference call sites src/audio_rag/retrieval/qdrant.py:285-288,
// 317-332; arithmetic in qdrant-client local/distances.py::cosine_similarity + np.argsort).
//
// Shape: persistent grid of one CTA per SM.  Warp 0 is the producer: one elected lane streams 16-row tiles
// (32 KB at 1024-d) of the bf16 corpus into a shared-memory ring with 1-D bulk async copies (TMA engine,
// cp.async.bulk -> SASS UBLKCP) completing on mbarriers.  Measured on B200: a 3-stage ring filled in 8 KB pieces
// (96 KB in flight per SM) reaches 6.7 TB/s, deeper rings are slower (6 stages: 6.2 TB/s).  Warps 1-8 consume: each
// takes two rows of a tile, reads them with conflict-free 128-bit LDS, multiplies against the query held in
// registers (fp32 FMA), butterfly-reduces, and appends scores above the running threshold to the CTA's candidate
// buffer in shared memory, which all consumer warps compact together (bitonic, keep the best Lc) when it nears
// capacity.  Thresholds are shared grid-wide through a monotone atomicMax so that after the first few tiles almost
// no row passes the compare.  The score vector never reaches HBM; each CTA emits one sorted list of Lc keys per query.
//
// Roofline: HBM.  Algorithmic bytes per launch = n_rows * dim * 2.
#include "common.cuh"
#include "engine.h"

namespace b200rag {

struct DenseScanParams {
    const uint16_t* corpus;
    int64_t n_rows;
    int64_t n_tiles;
    const uint16_t* q_bits;        // [NQ, dim]
    const uint32_t* masks[2];      // per query eligibility bitmap or nullptr
    uint64_t* g_thr;               // [NQ] grid-wide running thresholds
    uint64_t* out;                 // [NQ][grid][Lc]
    int64_t out_q_stride;          // grid * Lc
    int Lc, cap, stages;
    int interleave;
    int split;
    unsigned long long* tile_counter;   // dynamic tile assignment: next unclaimed tile (zeroed before the launch), or null
};

constexpr int kScanThreads = 32 + 32 * kScanConsumerWarps;

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// SH == 0: every consumer warp appends to its own shared-memory buffer and bitonic-compacts it when it fills (cheap for
//          small top-k: 64-key buffers).
// SH == 1: larger top-k: ONE buffer per CTA and query.  A warp-private 512-key compaction takes ~10 us during which the
//          warp's two rows of every tile are not consumed, so the whole ring stalls (measured: top-100 scans at 4.6-5.4
//          TB/s instead of 6.3).  The shared buffer sees 8x the rows, so its threshold is 8x tighter (5x fewer pushes),
//          and all 8 warps sort it together.  Consumers meet at a named barrier every kSyncTiles tiles; the decision to
//          compact is taken one interval ahead by one thread, so every warp sees the same decision.
constexpr int kSyncTiles = 8;
constexpr int kTileChunk = 8;           // tiles claimed per atomic in dynamic mode
constexpr int kSharedMargin = 2 * kSyncTiles * kScanTileRows;   // pushes possible between decision and compaction

template <int NCH, int NQ, int SH>
__global__ void __launch_bounds__(kScanThreads, 1) dense_scan_kernel(const DenseScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int T = kScanTileRows;
    constexpr int ROW_BYTES = NCH * 512;
    constexpr int STAGE_BYTES = T * ROW_BYTES;
    constexpr int DIM = NCH * 256;

    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * STAGE_BYTES);
    uint64_t* empty = full + p.stages;
    uint64_t* sthr = empty + p.stages;  // [stage][2] grid-wide thresholds sampled by the producer with each tile
    long long* stile = reinterpret_cast<long long*>(sthr + 2 * p.stages);   // [stage] the tile in the stage, -1 = no more tiles
    uint64_t* bufs = reinterpret_cast<uint64_t*>(stile + p.stages);  // SH == 0: [consumer warp][NQ][cap];  SH == 1: [NQ][cap] | cthr[NQ] | ccnt[NQ] | flag[NQ]
    uint64_t* cthr = bufs + (size_t)NQ * p.cap;
    int* ccnt = reinterpret_cast<int*>(cthr + NQ);
    int* cflag = ccnt + NQ;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kScanConsumerWarps);
        }
        fence_mbar_init();
        if (SH) {
            for (int q = 0; q < NQ; ++q) { cthr[q] = 0; ccnt[q] = 0; cflag[q] = 0; }
        }
    }
    __syncthreads();

    // tile order: interleaved (tile t -> CTA t mod grid: all SMs sweep one contiguous window of the corpus together)
    // or contiguous ranges per CTA
    const int64_t t0 = p.interleave ? (int64_t)blockIdx.x : p.n_tiles * (int64_t)blockIdx.x / (int64_t)gridDim.x;
    const int64_t t1 = p.interleave ? p.n_tiles : p.n_tiles * (int64_t)(blockIdx.x + 1) / (int64_t)gridDim.x;
    const int64_t tstep = p.interleave ? (int64_t)gridDim.x : 1;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            uint64_t gcur[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) gcur[q] = 0;
            // Tile order.  Dynamic (default): tiles are claimed kTileChunk at a time from a grid-wide counter, the claim for
            // the NEXT chunk issued when a chunk starts so that its latency hides behind the copies -- all CTAs still
            // sweep one contiguous window of the corpus together, but an SM that also hosts the sparse leg's or a tail's
            // CTAs simply takes fewer tiles instead of holding the whole scan back (static split, 12.5M rows, top-100,
            // pipelined: 4.36 ms; the tails of the previous search co-reside with the scan).  Static: tile t -> CTA
            // t mod grid (interleave) or contiguous ranges.
            const bool dyn = p.tile_counter != nullptr;
            long long cur = 0, cend = 0, nbase = 0;
            if (dyn) {
                cur = (long long)atomicAdd(p.tile_counter, (unsigned long long)kTileChunk);
                cend = cur + kTileChunk;
                nbase = (long long)atomicAdd(p.tile_counter, (unsigned long long)kTileChunk);
            } else {
                cur = t0;
            }
            for (;;) {
                long long t;
                if (dyn) {
                    if (cur >= cend) {
                        cur = nbase;
                        cend = cur + kTileChunk;
                        if (cur < p.n_tiles) nbase = (long long)atomicAdd(p.tile_counter, (unsigned long long)kTileChunk);
                    }
                    t = cur < p.n_tiles ? cur++ : -1;
                } else {
                    t = cur < t1 ? cur : -1;
                    cur += tstep;
                }
                mbar_wait(&empty[st], ph ^ 1u);
                stile[st] = t;
                if (t < 0) {                      // end marker: the stage completes on this arrive alone
                    mbar_arrive(&full[st]);
                    break;
                }
                const int64_t row0 = t * T;
                const int64_t left = p.n_rows - row0;
                const uint32_t rows = left < T ? (uint32_t)left : (uint32_t)T;
                const uint32_t bytes = rows * ROW_BYTES;
                // one volatile read of the grid-wide thresholds per tile per CTA, handed to the consumers with the
                // stage (ordered by the barrier arrive below), instead of one read per consumer warp
#pragma unroll
                for (int q = 0; q < NQ; ++q) sthr[st * 2 + q] = gcur[q];
                mbar_arrive_expect_tx(&full[st], bytes);
                {
                    // one stage = `split` bulk copies (all complete on the same barrier)
                    const uint32_t piece = (uint32_t)STAGE_BYTES / (uint32_t)p.split;
                    uint8_t* dst = ring + (size_t)st * STAGE_BYTES;
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.corpus + row0 * DIM);
                    for (uint32_t o = 0; o < bytes; o += piece)
                        bulk_g2s(dst + o, src + o, bytes - o < piece ? bytes - o : piece, &full[st]);
                }
                // sample for the NEXT tile now: the load's latency hides behind the wait for a free stage
#pragma unroll
                for (int q = 0; q < NQ; ++q) gcur[q] = ld_volatile_u64(&p.g_thr[q]);
                if (++st == p.stages) { st = 0; ph ^= 1u; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int cw = warp - 1;
    const int ctid = threadIdx.x - 32;
    constexpr int NCT = 32 * kScanConsumerWarps;

    float qf[NQ][NCH * 8];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const uint4 v = *reinterpret_cast<const uint4*>(p.q_bits + (size_t)q * DIM + j * 256 + lane * 8);
            qf[q][j * 8 + 0] = bf16lo(v.x); qf[q][j * 8 + 1] = bf16hi(v.x);
            qf[q][j * 8 + 2] = bf16lo(v.y); qf[q][j * 8 + 3] = bf16hi(v.y);
            qf[q][j * 8 + 4] = bf16lo(v.z); qf[q][j * 8 + 5] = bf16hi(v.z);
            qf[q][j * 8 + 6] = bf16lo(v.w); qf[q][j * 8 + 7] = bf16hi(v.w);
        }

    uint64_t thr[NQ];
    int cnt[NQ];
    uint64_t* mybuf[NQ];
    int tcount = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        thr[q] = 0;
        cnt[q] = 0;
        mybuf[q] = SH ? bufs + (size_t)q * p.cap : bufs + ((size_t)cw * NQ + q) * p.cap;
    }

    // SH: sort the CTA's buffer of query q (all consumer warps), keep the best Lc, publish the new threshold
    auto compact_shared = [&](int q) {
        const int n = ccnt[q];
        const int np2 = next_pow2(n > p.Lc ? n : p.Lc);     // sort only what is there (the final compaction is short)
        named_bar_sync(1, NCT);
        for (int i = n + ctid; i < np2; i += NCT) mybuf[q][i] = 0;
        cta_bitonic_desc(mybuf[q], np2, ctid, NCT, 1);
        if (ctid == 0) {
            if (n > p.Lc) ccnt[q] = p.Lc;
            const uint64_t nt = mybuf[q][p.Lc - 1];
            if (nt > cthr[q]) {
                cthr[q] = nt;
                atomicMax(reinterpret_cast<unsigned long long*>(&p.g_thr[q]), (unsigned long long)nt);
            }
        }
        named_bar_sync(1, NCT);
    };

    int st = 0;
    uint32_t ph = 0;
    for (;;) {
        mbar_wait(&full[st], ph);
        const long long t = stile[st];
        if (t < 0) break;                 // (every consumer warp sees the same sequence of tiles and the same end)
        // refresh from the grid-wide thresholds (monotone; any warp's Lc-th best is a valid lower bound)
        uint64_t g[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) g[q] = sthr[st * 2 + q];
        const uint8_t* sp = ring + (size_t)st * STAGE_BYTES;
        const int64_t row0 = t * T;
        const int64_t left = p.n_rows - row0;
        const int rows = left < T ? (int)left : T;
        const bool has0 = cw < rows, has1 = (cw + kScanConsumerWarps) < rows;

        uint4 v0[NCH], v1[NCH];
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            v0[j] = has0 ? *reinterpret_cast<const uint4*>(sp + cw * ROW_BYTES + j * 512 + lane * 16)
                         : make_uint4(0, 0, 0, 0);
            v1[j] = has1 ? *reinterpret_cast<const uint4*>(sp + (cw + kScanConsumerWarps) * ROW_BYTES + j * 512 +
                                                           lane * 16)
                         : make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == p.stages) { st = 0; ph ^= 1u; }

        float tot[2][NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const float* qq = &qf[q][j * 8];
                a0 = fmaf(bf16lo(v0[j].x), qq[0], a0); b0 = fmaf(bf16hi(v0[j].x), qq[1], b0);
                a0 = fmaf(bf16lo(v0[j].y), qq[2], a0); b0 = fmaf(bf16hi(v0[j].y), qq[3], b0);
                a0 = fmaf(bf16lo(v0[j].z), qq[4], a0); b0 = fmaf(bf16hi(v0[j].z), qq[5], b0);
                a0 = fmaf(bf16lo(v0[j].w), qq[6], a0); b0 = fmaf(bf16hi(v0[j].w), qq[7], b0);
                a1 = fmaf(bf16lo(v1[j].x), qq[0], a1); b1 = fmaf(bf16hi(v1[j].x), qq[1], b1);
                a1 = fmaf(bf16lo(v1[j].y), qq[2], a1); b1 = fmaf(bf16hi(v1[j].y), qq[3], b1);
                a1 = fmaf(bf16lo(v1[j].z), qq[4], a1); b1 = fmaf(bf16hi(v1[j].z), qq[5], b1);
                a1 = fmaf(bf16lo(v1[j].w), qq[6], a1); b1 = fmaf(bf16hi(v1[j].w), qq[7], b1);
            }
            float s0 = a0 + b0, s1 = a1 + b1;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, d);
                s1 += __shfl_xor_sync(0xffffffffu, s1, d);
            }
            tot[0][q] = s0;
            tot[1][q] = s1;
        }

#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (g[q] > thr[q]) thr[q] = g[q];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const bool has = r == 0 ? has0 : has1;
                if (!has) continue;
                const uint32_t row = (uint32_t)(row0 + cw + r * kScanConsumerWarps);
                const uint64_t key = make_key(tot[r][q] + 0.0f, row);
                if (key > thr[q]) {  // warp-uniform: every lane holds the same reduced score
                    bool ok = true;
                    const uint32_t* m = p.masks[q];
                    if (m != nullptr) ok = (m[row >> 5] >> (row & 31)) & 1u;
                    if (ok && SH) {
                        if (lane == 0) {
                            const int pos = atomicAdd(&ccnt[q], 1);   // room is guaranteed by the margin (see below)
                            mybuf[q][pos] = key;
                        }
                    } else if (ok) {
                        if (lane == 0) mybuf[q][cnt[q]] = key;
                        if (++cnt[q] == p.cap) {
                            warp_bitonic_desc(mybuf[q], p.cap, lane);
                            cnt[q] = p.Lc;
                            const uint64_t nt = mybuf[q][p.Lc - 1];
                            if (nt > thr[q]) thr[q] = nt;
                            if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(&p.g_thr[q]),
                                                     (unsigned long long)nt);
                            __syncwarp();
                        }
                    }
                }
            }
        }
        if (SH && (++tcount & (kSyncTiles - 1)) == 0) {
            // Every consumer warp has processed the same tiles.  cflag was written by one thread BEFORE this barrier (at
            // the previous meeting), so all warps read the same decision; pushes that race with the decision are
            // covered by the margin: a buffer is compacted while it still has >= kSharedMargin / 2 free slots.
            named_bar_sync(1, NCT);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (cflag[q]) compact_shared(q);
                const uint64_t ct = cthr[q];
                if (ct > thr[q]) thr[q] = ct;
            }
            named_bar_sync(1, NCT);
            if (ctid == 0) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) cflag[q] = ccnt[q] > p.cap - kSharedMargin ? 1 : 0;
            }
        }
    }

    if constexpr (SH != 0) {
        // one sorted list per query straight from the CTA's buffer
        named_bar_sync(1, NCT);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            compact_shared(q);
            uint64_t* o = p.out + (size_t)q * p.out_q_stride + (size_t)blockIdx.x * p.Lc;
            for (int i = ctid; i < p.Lc; i += NCT) o[i] = mybuf[q][i];
        }
    } else {
    // ---- per-warp final compaction, then one CTA-level merge per query (ring memory is free now)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        __syncwarp();
        for (int i = cnt[q] + lane; i < p.cap; i += 32) mybuf[q][i] = 0;
        warp_bitonic_desc(mybuf[q], p.cap, lane);
    }
    uint64_t* marea = reinterpret_cast<uint64_t*>(ring);
    const int mcount = kScanConsumerWarps * p.Lc;
    const int mpow2 = next_pow2(mcount);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        named_bar_sync(1, NCT);  // every warp is past its last ring read / previous query's output
        for (int i = lane; i < p.Lc; i += 32) marea[cw * p.Lc + i] = mybuf[q][i];
        for (int i = mcount + ctid; i < mpow2; i += NCT) marea[i] = 0;
        cta_bitonic_desc(marea, mpow2, ctid, NCT, 1);
        uint64_t* o = p.out + (size_t)q * p.out_q_stride + (size_t)blockIdx.x * p.Lc;
        for (int i = ctid; i < p.Lc; i += NCT) o[i] = marea[i];
        // grid-wide threshold gets the CTA's Lc-th best as well
        if (ctid == 0 && marea[p.Lc - 1] != 0)
            atomicMax(reinterpret_cast<unsigned long long*>(&p.g_thr[q]), (unsigned long long)marea[p.Lc - 1]);
    }
    }
}

template <int NCH, int NQ, int SH>
static int launch_one(Shard* s, const DenseScanParams& p, int grid, size_t smem) {
    auto kern = dense_scan_kernel<NCH, NQ, SH>;
    // attributes are set once per instantiation (and again only if a larger footprint is needed): two driver calls
    // fewer between the start of a search and its first kernel
    static AttrCache attr;
    if (attr.raise(s->cfg.device, smem)) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // keep the SM's shared-memory carve-out at its maximum: the ring only needs ~105 KB, and the sparse leg's CTAs
    // (side stream) are meant to co-reside in what is left
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    kern<<<grid, kScanThreads, smem, s->stream>>>(p);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

int dense_scan_nlists(const Shard* s) { return (s->scan_ctas > 0 && s->scan_ctas < s->sm_count) ? s->scan_ctas : s->sm_count; }

int launch_dense_scan(Shard* s, int batch, int Lc, uint64_t* out_lists, int* nlists) {
    const int nch = s->dim / 256;
    const int64_t n_tiles = (s->n_rows + kScanTileRows - 1) / kScanTileRows;
    int grid = dense_scan_nlists(s);
    if (n_tiles < grid) grid = (int)(n_tiles > 0 ? n_tiles : 1);
    *nlists = grid;

    // selection flavour: one CTA-shared buffer (default) or warp-private buffers (knob: B200RAG_SCAN_SHARED=0)
    const bool sh = s->scan_shared != 0;   // measured on B200 (10M rows): shared beats warp-private at every top-k (top-10: 3.18 vs 3.30 ms, top-100: 3.25 vs 3.95 ms)
    const int cap = sh ? next_pow2(Lc + kSharedMargin + 64) : (next_pow2(Lc + 32) < 64 ? 64 : next_pow2(Lc + 32));
    const size_t stage_bytes = (size_t)kScanTileRows * nch * 512;
    const size_t max_smem = 227 * 1024;
    const size_t merge_bytes = sh ? 0 : (size_t)next_pow2(kScanConsumerWarps * Lc) * 8;

    s->stats.dense_path = 1;
    s->stats.dense_passes = 0;

    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[0], s->stream)); }
    int q = 0;
    while (q < batch) {
        int nq = (batch - q >= 2) ? 2 : 1;
        size_t buf_bytes = sh ? (size_t)nq * cap * 8 + 64 : (size_t)kScanConsumerWarps * nq * cap * 8;
        if (nq == 2 && buf_bytes + 2 * stage_bytes + 256 > max_smem) {
            nq = 1;
            buf_bytes = sh ? (size_t)cap * 8 + 64 : (size_t)kScanConsumerWarps * cap * 8;
        }
        int stages = (int)((max_smem - buf_bytes - 512) / stage_bytes);
        if (stages > 8) stages = 8;
        if (s->dense_stage_cap > 0 && stages > s->dense_stage_cap) stages = s->dense_stage_cap;
        while ((size_t)stages * stage_bytes < merge_bytes) ++stages;
        if (stages < 2) { set_error("dense_scan: top-k too large for shared memory"); return B200RAG_ERR_INVALID; }
        const size_t smem = (size_t)stages * stage_bytes + (size_t)stages * 40 + buf_bytes;
        if (smem > max_smem) { set_error("dense_scan: shared memory budget exceeded"); return B200RAG_ERR_INVALID; }

        DenseScanParams p{};
        p.corpus = s->dense.as<uint16_t>();
        p.n_rows = s->n_rows;
        p.n_tiles = n_tiles;
        p.q_bits = s->ws.q_bits.as<uint16_t>() + (size_t)q * s->dim;
        p.masks[0] = s->h_masks.empty() ? nullptr : s->h_masks[q];
        p.masks[1] = (nq == 2 && !s->h_masks.empty()) ? s->h_masks[q + 1] : nullptr;
        p.g_thr = s->ws.thr.as<uint64_t>() + (size_t)s->thr_par * 2 * batch + q;
        p.out = out_lists + (size_t)q * grid * Lc;
        p.out_q_stride = (int64_t)grid * Lc;
        p.Lc = Lc;
        p.cap = cap;
        p.stages = stages;
        p.interleave = s->tile_interleave;
        p.split = s->bulk_split;
        // dynamic tile assignment: one counter per pass, zeroed by the caller's threshold memset
        // (ws.thr: per call parity [B thresholds | B tile counters], then the postings counter)
        p.tile_counter = s->scan_dynamic ? reinterpret_cast<unsigned long long*>(s->ws.thr.as<uint64_t>() + (size_t)s->thr_par * 2 * batch + batch + q) : nullptr;

        int rc;
#define B2_SCAN_CASE(NCH_, NQ_) (sh ? launch_one<NCH_, NQ_, 1>(s, p, grid, smem) : launch_one<NCH_, NQ_, 0>(s, p, grid, smem))
        if (nq == 2) {
            switch (nch) {
                case 1: rc = B2_SCAN_CASE(1, 2); break;
                case 2: rc = B2_SCAN_CASE(2, 2); break;
                case 3: rc = B2_SCAN_CASE(3, 2); break;
                default: rc = B2_SCAN_CASE(4, 2); break;
            }
        } else {
            switch (nch) {
                case 1: rc = B2_SCAN_CASE(1, 1); break;
                case 2: rc = B2_SCAN_CASE(2, 1); break;
                case 3: rc = B2_SCAN_CASE(3, 1); break;
                default: rc = B2_SCAN_CASE(4, 1); break;
            }
        }
#undef B2_SCAN_CASE
        if (rc != B200RAG_OK) return rc;
        s->stats.dense_passes++;
        q += nq;
    }
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[1], s->stream)); s->ev_dense = true; }
    s->stats.dense_bytes = (int64_t)s->stats.dense_passes * s->n_rows * s->dim * 2;
    return B200RAG_OK;
}

}  // namespace b200rag
