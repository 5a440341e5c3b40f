// K1c  dense_scan_q8: the bandwidth-bound dense scan over an 8-bit COPY of the corpus ("optional
// fp8/int8 corpus compression with exact re-score of the top candidates").  Opt-in (b200rag_set_compression).
//
// Replaces, like dense_scan.cu, the dense leg of client.query_points (src/audio_rag/retrieval/qdrant.py:285-288, 317-332)
// for 1-2 queries per corpus pass.  The result is still EXACT: this kernel only selects candidates; they are re-scored
// from the bf16 rows in the canonical order (select.cu) and the guard below is rigorous.
//
// Format: row r = 1024 (dim) int8 values q_i + trailer {f32 scale, f32 l1, f32 e2, 4 bytes pad}: x_i ~ scale * q_i with
// |x_i - scale * q_i| <= 0.5 * scale (symmetric, round to nearest, scale = max|x_i| / 127), l1 = scale * sum|q_i|,
// e2 = ||x - scale * q||_2, the row's MEASURED quantisation residual.
// The query is quantised to 14-bit integers in the kernel's prologue (qs = max|y_i| / 8191), so a row's score is ONE
// exact integer dot product (dp2a: 16-bit x 8-bit multiply-accumulates, |sum| <= 1024 * 8191 * 127 < 2^31):
//     s^ = scale * qs * sum_i q_i * yq_i
//     s - s^ = sum_i (x_i - scale q_i) y_i  +  sum_i scale q_i (y_i - qs yq_i)
//     |s - s^| <= min(0.5 * scale * ||y||_1, e2 * ||y||_2)  +  0.5 * qs * l1  (+ fp32 rounding of the products)
// (Hoelder and Cauchy-Schwarz on the first sum; the second is the tighter one on typical rows -- e2 ~ scale * sqrt(dim / 12)
//  against 0.5 * scale * ||y||_1 ~ 0.4 * scale * sqrt(dim) -- and the band it leaves holds about half as many rows.)
// The kernel ranks rows by the UPPER BOUND  ub = s^ + err(row): a row that is not among the Lc retained candidates has
// exact score <= ub <= the weakest retained ub, so the leg is exact as soon as that weakest ub is below the L-th exact
// score -- the same guard as for the bf16 scan, with eps = 0 and a larger slack (typically ~250 rows fall inside the
// error band at 10M rows).  When the guard does not clear, the retry takes the bf16 scan.
//
// Shape: the bf16 scan's (dense_scan.cu): persistent grid, one CTA per SM, a producer lane streaming 32-row tiles
// (33 280 B) through a 3-stage ring with bulk async copies on mbarriers, 8 consumer warps, one CTA-shared candidate
// buffer per query, grid-wide thresholds through a monotone atomicMax.
//
// Roofline: HBM.  Algorithmic bytes per launch = n_rows * (dim + 16)  -- half the bf16 scan's.
#include "common.cuh"
#include "engine.h"

namespace b200rag {

constexpr int kQ8Trailer = 16;
constexpr int kQ8TileRows = 32;
constexpr int kQ8SyncTiles = 8;
constexpr int kQ8Margin = 2 * kQ8SyncTiles * kQ8TileRows;     // pushes possible between a compaction decision and the compaction
constexpr int kQ8Threads = 32 + 32 * kScanConsumerWarps;
constexpr float kQ8Half = 0.5005f;                            // 0.5 + slack for the fp32 rounding of x / scale at build time

struct DenseQ8Params {
    const uint8_t* corpus;         // [n_rows][dim + 16]
    int64_t n_rows;
    int64_t n_tiles;
    const uint16_t* q_bits;        // [NQ, dim] bf16 bits of the unit queries
    const uint32_t* masks[2];
    uint64_t* g_thr;               // [NQ] grid-wide running thresholds
    uint64_t* out;                 // [NQ][grid][Lc]
    int64_t out_q_stride;
    int Lc, cap, stages, split;
};

// ---------------------------------------------------------------------------------------------- quantiser
// one warp per row: bf16 row -> int8 row + {scale, l1}
__global__ void __launch_bounds__(256) quantize_rows_kernel(const uint16_t* __restrict__ dense, int64_t row0, int64_t n,
                                                            int dim, uint8_t* __restrict__ q8) {
    const int64_t r = row0 + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= row0 + n) return;
    const uint16_t* src = dense + (size_t)r * dim;
    uint8_t* dst = q8 + (size_t)r * (dim + kQ8Trailer);
    float m = 0.f;
    for (int k = lane; k < dim; k += 32) m = fmaxf(m, fabsf(__uint_as_float((uint32_t)src[k] << 16)));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    const float scale = m > 0.f ? m / 127.f : 1.f;
    const float inv = 1.f / scale;
    int l1 = 0;
    float e2 = 0.f;
    for (int k = lane; k < dim; k += 32) {
        const float x = __uint_as_float((uint32_t)src[k] << 16);
        int v = __float2int_rn(x * inv);
        v = v > 127 ? 127 : (v < -127 ? -127 : v);
        dst[k] = (uint8_t)(int8_t)v;
        l1 += v < 0 ? -v : v;
        const float e = fmaf(-scale, (float)v, x);      // x - scale * v, rounded once
        e2 = fmaf(e, e, e2);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, d);
        e2 += __shfl_xor_sync(0xffffffffu, e2, d);
    }
    if (lane == 0) {
        float* tr = reinterpret_cast<float*>(dst + dim);
        tr[0] = scale;
        tr[1] = scale * (float)l1 * 1.0001f;
        tr[2] = sqrtf(e2) * 1.0001f;                     // (1e-4 covers the fp32 roundings of the residuals and their sum)
        tr[3] = 0.f;
    }
}

int launch_quantize_rows(Shard* s, int64_t row0, int64_t n) {
    if (n <= 0) return B200RAG_OK;
    const int64_t threads = n * 32;
    quantize_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s->stream>>>(s->dense.as<uint16_t>(), row0, n, s->dim,
                                                                                 s->dense_q8.as<uint8_t>());
    B2_CUDA(cudaGetLastError());
    return B200RAG_OK;
}

// ---------------------------------------------------------------------------------------------- scan
// NW16 = dim / 64: number of 16-byte chunks a lane reads per row (dim / 32 bytes per lane)
template <int NW16, int NQ>
__global__ void __launch_bounds__(kQ8Threads, 1) dense_scan_q8_kernel(const DenseQ8Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int T = kQ8TileRows;
    constexpr int DIM = NW16 * 512;                   // NW16 chunks x 32 lanes x 16 bytes
    constexpr int ROW_BYTES = DIM + kQ8Trailer;
    constexpr int STAGE_BYTES = T * ROW_BYTES;
    constexpr int ROWS_PER_WARP = T / kScanConsumerWarps;

    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * STAGE_BYTES);
    uint64_t* empty = full + p.stages;
    uint64_t* sthr = empty + p.stages;               // [stage][2]
    uint64_t* bufs = sthr + 2 * p.stages;            // [NQ][cap] | cthr[NQ] | ccnt[NQ] | cflag[NQ]
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
        for (int q = 0; q < NQ; ++q) { cthr[q] = 0; ccnt[q] = 0; cflag[q] = 0; }
    }
    __syncthreads();

    const int64_t t0 = blockIdx.x, t1 = p.n_tiles, tstep = gridDim.x;      // interleaved: tile t -> CTA t mod grid

    if (warp == 0) {
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            uint64_t gcur[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) gcur[q] = 0;
            for (int64_t t = t0; t < t1; t += tstep) {
                mbar_wait(&empty[st], ph ^ 1u);
                const int64_t row0 = t * T;
                const int64_t left = p.n_rows - row0;
                const uint32_t rows = left < T ? (uint32_t)left : (uint32_t)T;
                const uint32_t bytes = rows * ROW_BYTES;
#pragma unroll
                for (int q = 0; q < NQ; ++q) sthr[st * 2 + q] = gcur[q];
                mbar_arrive_expect_tx(&full[st], bytes);
                {
                    const uint32_t piece = (uint32_t)STAGE_BYTES / (uint32_t)p.split;
                    uint8_t* dst = ring + (size_t)st * STAGE_BYTES;
                    const uint8_t* src = p.corpus + (size_t)row0 * ROW_BYTES;
                    for (uint32_t o = 0; o < bytes; o += piece)
                        bulk_g2s(dst + o, src + o, bytes - o < piece ? bytes - o : piece, &full[st]);
                }
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

    // query -> 14-bit integers, element pairs packed for dp2a: lane owns bytes [c * 512 + lane * 16, +16) of a row, c < NW16
    int qp[NQ][NW16 * 8];            // 16 elements per chunk = 8 packed pairs
    float qs[NQ], ql1[NQ], ql2[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float y[NW16 * 16];
        float m = 0.f, l1 = 0.f, l2 = 0.f;
#pragma unroll
        for (int c = 0; c < NW16; ++c) {
            const uint4 a = *reinterpret_cast<const uint4*>(p.q_bits + (size_t)q * DIM + c * 512 + lane * 16);
            const uint4 b = *reinterpret_cast<const uint4*>(p.q_bits + (size_t)q * DIM + c * 512 + lane * 16 + 8);
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                y[c * 16 + 2 * i] = __uint_as_float(w[i] << 16);
                y[c * 16 + 2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
            }
        }
#pragma unroll
        for (int i = 0; i < NW16 * 16; ++i) { m = fmaxf(m, fabsf(y[i])); l1 += fabsf(y[i]); l2 = fmaf(y[i], y[i], l2); }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
            l1 += __shfl_xor_sync(0xffffffffu, l1, d);
            l2 += __shfl_xor_sync(0xffffffffu, l2, d);
        }
        const float sc = m > 0.f ? m / 8191.f : 1.f;
        const float inv = 1.f / sc;
        qs[q] = sc;
        ql1[q] = l1 * 1.0001f;                        // (the fp32 sum of 1024 magnitudes: 1e-4 covers its rounding)
        ql2[q] = sqrtf(l2) * 1.0001f;
#pragma unroll
        for (int i = 0; i < NW16 * 8; ++i) {
            int lo = __float2int_rn(y[2 * i] * inv), hi = __float2int_rn(y[2 * i + 1] * inv);
            lo = lo > 8191 ? 8191 : (lo < -8191 ? -8191 : lo);
            hi = hi > 8191 ? 8191 : (hi < -8191 ? -8191 : hi);
            qp[q][i] = (int)(((uint32_t)hi << 16) | ((uint32_t)lo & 0xFFFFu));
        }
    }

    uint64_t thr[NQ];
    uint64_t* mybuf[NQ];
    int tcount = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) { thr[q] = 0; mybuf[q] = bufs + (size_t)q * p.cap; }

    auto compact_shared = [&](int q) {
        const int n = ccnt[q];
        const int np2 = next_pow2(n > p.Lc ? n : p.Lc);
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
    for (int64_t t = t0; t < t1; t += tstep) {
        mbar_wait(&full[st], ph);
        uint64_t g[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) g[q] = sthr[st * 2 + q];
        const uint8_t* sp = ring + (size_t)st * STAGE_BYTES;
        const int64_t row0 = t * T;
        const int64_t left = p.n_rows - row0;
        const int rows = left < T ? (int)left : T;

        // this warp's rows: cw, cw + 8, cw + 16, cw + 24
        uint4 v[ROWS_PER_WARP][NW16];
        float rscale[ROWS_PER_WARP], rl1[ROWS_PER_WARP], re2[ROWS_PER_WARP];
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; ++r) {
            const int rr = cw + r * kScanConsumerWarps;
            const bool has = rr < rows;
            const uint8_t* rp = sp + (size_t)rr * ROW_BYTES;
#pragma unroll
            for (int c = 0; c < NW16; ++c)
                v[r][c] = has ? *reinterpret_cast<const uint4*>(rp + c * 512 + lane * 16) : make_uint4(0, 0, 0, 0);
            const float4 tr = has ? *reinterpret_cast<const float4*>(rp + DIM) : make_float4(0.f, 0.f, 0.f, 0.f);
            rscale[r] = tr.x;
            rl1[r] = tr.y;
            re2[r] = tr.z;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == p.stages) { st = 0; ph ^= 1u; }

        // integer dot products of the warp's rows with every query: independent accumulator chains (two per row and
        // query), all of them reduced across the lanes before any branch, so that the chains and the shuffles of the
        // 4 rows interleave
        int dot[NQ][ROWS_PER_WARP];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int r = 0; r < ROWS_PER_WARP; ++r) {
                int a0 = 0, a1 = 0;
#pragma unroll
                for (int c = 0; c < NW16; ++c) {
                    const uint32_t w[4] = {v[r][c].x, v[r][c].y, v[r][c].z, v[r][c].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        a0 = __dp2a_lo(qp[q][c * 8 + 2 * i], (int)w[i], a0);
                        a1 = __dp2a_hi(qp[q][c * 8 + 2 * i + 1], (int)w[i], a1);
                    }
                }
                dot[q][r] = a0 + a1;
            }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int r = 0; r < ROWS_PER_WARP; ++r) dot[q][r] += __shfl_xor_sync(0xffffffffu, dot[q][r], d);

#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (g[q] > thr[q]) thr[q] = g[q];
#pragma unroll
            for (int r = 0; r < ROWS_PER_WARP; ++r) {
                const int rr = cw + r * kScanConsumerWarps;
                if (rr >= rows) continue;
                const float sh = rscale[r] * qs[q] * (float)dot[q][r];
                // rigorous upper bound of the exact score (see the header); fp32 slack on top
                const float erow = fminf(kQ8Half * rscale[r] * ql1[q], re2[r] * ql2[q]);
                const float ub = sh + erow + kQ8Half * qs[q] * rl1[r] + 4e-7f * fabsf(sh) + 1e-7f;
                const uint32_t row = (uint32_t)(row0 + rr);
                const uint64_t key = make_key(ub + 0.0f, row);
                if (key > thr[q]) {                   // warp-uniform
                    bool ok = true;
                    const uint32_t* m = p.masks[q];
                    if (m != nullptr) ok = (m[row >> 5] >> (row & 31)) & 1u;
                    if (ok && lane == 0) {
                        const int pos = atomicAdd(&ccnt[q], 1);
                        mybuf[q][pos] = key;
                    }
                }
            }
        }
        if ((++tcount & (kQ8SyncTiles - 1)) == 0) {
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
                for (int q = 0; q < NQ; ++q) cflag[q] = ccnt[q] > p.cap - kQ8Margin ? 1 : 0;
            }
        }
    }

    named_bar_sync(1, NCT);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        compact_shared(q);
        uint64_t* o = p.out + (size_t)q * p.out_q_stride + (size_t)blockIdx.x * p.Lc;
        for (int i = ctid; i < p.Lc; i += NCT) o[i] = mybuf[q][i];
    }
}

template <int NW16, int NQ>
static int launch_q8_one(Shard* s, const DenseQ8Params& p, int grid, size_t smem) {
    auto kern = dense_scan_q8_kernel<NW16, NQ>;
    static AttrCache attr;
    if (attr.raise(s->cfg.device, smem)) {
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    kern<<<grid, kQ8Threads, smem, s->stream>>>(p);
    B2_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return B200RAG_OK;
}

bool dense_q8_supported(const Shard* s) { return s->dim == 512 || s->dim == 1024; }

int launch_dense_scan_q8(Shard* s, int batch, int Lc, uint64_t* out_lists, int* nlists) {
    const int64_t n_tiles = (s->n_rows + kQ8TileRows - 1) / kQ8TileRows;
    int grid = dense_scan_nlists(s);
    if (n_tiles < grid) grid = (int)(n_tiles > 0 ? n_tiles : 1);
    *nlists = grid;
    const int cap = next_pow2(Lc + kQ8Margin + 64);
    const size_t stage_bytes = (size_t)kQ8TileRows * (s->dim + kQ8Trailer);
    const size_t max_smem = 227 * 1024;
    s->stats.dense_path = 3;
    s->stats.dense_passes = 0;
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[0], s->stream)); }
    int q = 0;
    while (q < batch) {
        int nq = (batch - q >= 2) ? 2 : 1;
        size_t buf_bytes = (size_t)nq * cap * 8 + 64;
        if (nq == 2 && buf_bytes + 2 * stage_bytes + 512 > max_smem) { nq = 1; buf_bytes = (size_t)cap * 8 + 64; }
        int stages = (int)((max_smem - buf_bytes - 512) / stage_bytes);
        if (stages > 8) stages = 8;
        if (s->dense_stage_cap > 0 && stages > s->dense_stage_cap) stages = s->dense_stage_cap;
        if (stages < 2) { set_error("dense_scan_q8: top-k too large for shared memory"); return B200RAG_ERR_INVALID; }
        const size_t smem = (size_t)stages * stage_bytes + (size_t)stages * 32 + buf_bytes;
        DenseQ8Params p{};
        p.corpus = s->dense_q8.as<uint8_t>();
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
        p.split = s->bulk_split;
        int rc;
        if (s->dim == 1024) rc = nq == 2 ? launch_q8_one<2, 2>(s, p, grid, smem) : launch_q8_one<2, 1>(s, p, grid, smem);
        else rc = nq == 2 ? launch_q8_one<1, 2>(s, p, grid, smem) : launch_q8_one<1, 1>(s, p, grid, smem);
        if (rc != B200RAG_OK) return rc;
        s->stats.dense_passes++;
        q += nq;
    }
    if (s->profile) { B2_CUDA(cudaEventRecord(s->ev[1], s->stream)); s->ev_dense = true; }
    s->stats.dense_bytes = (int64_t)s->stats.dense_passes * s->n_rows * (s->dim + kQ8Trailer);
    return B200RAG_OK;
}

}  // namespace b200rag
