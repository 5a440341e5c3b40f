// Internal layout of a shard and the launchers each .cu exports to engine.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/b200rag.h"

namespace b200rag {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define B2_CUDA(call)                                             \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);     \
    } while (0)
#define B2_TRY(call)                       \
    do {                                   \
        int rc__ = (call);                 \
        if (rc__ != B200RAG_OK) return rc__; \
    } while (0)

// Growable device array (geometric growth, contents preserved).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;  // bytes
    int ensure(size_t bytes, size_t keep_bytes, cudaStream_t st);
    void release();
    template <typename T>
    T* as() const { return (T*)p; }
};

// Kernel function attributes (dynamic shared memory limit, carve-out) are per DEVICE: launchers cache what they set in
// a table indexed by the device ordinal, so a process that owns shards on several GPUs configures each of them.
constexpr int kMaxDevices = 64;
struct AttrCache {
    size_t smem[kMaxDevices] = {};
    // true when `bytes` exceeds what was configured on `device` so far (and records it)
    bool raise(int device, size_t bytes) {
        if (device < 0 || device >= kMaxDevices) return true;
        if (bytes <= smem[device]) return false;
        smem[device] = bytes;
        return true;
    }
};

constexpr int kMaxBatchMasks = 4096;
constexpr int kScanConsumerWarps = 8;
constexpr int kScanTileRows = 16;
constexpr int kSparseThreads = 256;
constexpr int kMaxQueryTermsChunk = 256;
constexpr int kMergeMaxKeys = 8192;    // smem bound of one merge group
constexpr int kMergeGroupKeys = 1024;  // preferred group size (keeps the bitonic network short)
constexpr int kTailMaxKeys = 1 << 20;  // most candidate keys (n_lists * Lc) the fused leg tail takes in one launch

struct DevPtr {
    void* p = nullptr;
    template <typename T>
    T* as() const { return (T*)p; }
};

struct Workspace {
                          // (the staged query blocks themselves live in Shard::slots: one H2D per batch:
                          //  [q_bits | q_sp_indptr | q_sp_terms | q_sp_w | q_masks])
    DevPtr q_bits;        // [B, dim] u16
    DevPtr q_sp_indptr;   // [B+1] i64
    DevPtr q_sp_terms;    // u32
    DevPtr q_sp_w;        // f32
    DevPtr q_masks;       // [B] const uint32_t*
    DevBuf thr;           // [B] u64 running grid-wide thresholds (dense) + [1] u64 postings counter
    DevPtr post_count;    // -> thr[B]
    DevBuf lists_a;       // candidate key lists (ping)
    DevBuf lists_b;       // candidate key lists (pong)
    DevBuf exact;         // [B, Lc] u64 exact keys
    DevBuf lists_c;       // second set for the sparse leg, which runs concurrently on the side stream
    DevBuf lists_d;
    DevBuf exact2;
    DevBuf q_eps;         // [B] f32 absolute error bound of the sparse scan's approximate scores | [B] i32 grid-wide thresholds
    DevBuf flag;          // 4-byte device flag for validation kernels
    DevBuf xpeers_dev;    // [world] void*: exchange windows of all ranks (peer pointers)
    DevBuf pool;          // filter path of the tcgen05 kernel: [B, cap] u64 keys | [B] i32 counters
    DevBuf cands;         // [nlegs, B, L] b200rag_cand  (single-shard search)
    DevBuf out;           // [B*top_k i64 ids | B*top_k f64 scores | B+1 i32 counts, ambiguous flag]
    DevBuf ex_keys;       // exhaustive path: [n_rows] u64 keys (all eligible rows -> exact keys)
    DevBuf ex_sorted;     // exhaustive path: [n_rows] u64 sorted keys
    DevBuf ex_temp;       // exhaustive path: cub temporary storage
};

// A staged query batch: its device block and the host-side facts `legs`/`fuse` need.  b200rag_stage fills slot 0;
// b200rag_stage_slot / b200rag_use_slot keep several batches resident so that searches can be enqueued back to back.
struct QuerySlot {
    DevBuf buf;
    b200rag_query q{};
    std::vector<const uint32_t*> h_masks;
    int64_t q_terms = 0;
    void *bits = nullptr, *ind = nullptr, *terms = nullptr, *w = nullptr, *masks = nullptr;
    bool staged = false;
};

constexpr int kProfileRing = 64;   // legs calls whose event timings stay readable (b200rag_get_stats_step)

struct Shard {
    b200rag_config cfg{};
    int dim = 0, vocab = 0, R = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;   // sparse leg of a hybrid search overlaps the dense scan here
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap_legs = true;
    bool overlap_force = false;
    int64_t overlap_max_rows = 0;         // overlap only shards up to this many rows (0 = always); tuning knobs via env
    int dense_stage_cap_env = 3;          // measured on B200: 3 stages x 32 KB in 8 KB pieces beats deeper rings (6.7 vs 6.2 TB/s)
    int tile_interleave = 1;
    int bulk_split = 4;
    int sparse_bpc = 0;                   // knob: blocks per sparse-scan CTA (0 = auto)
    int sparse_threads = 128;             // knob: threads per sparse-scan CTA (128 or 256)
    bool fused_tail = true;               // knob: merge + re-score + finalize of a leg in one launch when the lists fit shared memory
    size_t gemm_smem_reserve = 0;         // set while a hybrid batch overlaps its legs: smem the tcgen05 kernel leaves free per SM
    int overlap_gemm = 1;                 // knob: overlap the legs of BATCHED hybrid searches too (filter epilogue + fewer stages)
    bool gemm_pairs = true;               // knob: CTA pairs (cta_group::2, 256 queries per corpus pass) when > 128 queries remain
    bool gemm_l2_prefetch = false;        // knob (B200RAG_GEMM_L2_PREFETCH): corpus tiles prefetched into L2 one tile ahead
                                          // (measured SLOWER: 10M rows, B = 256 5.56 vs 4.79 ms, B = 128 5.78 vs 3.06 ms -- off)
    int gemm_stage_cap = 0;               // knob (B200RAG_GEMM_STAGES): cap on the tcgen05 kernel's pipeline depth (0 = as many as fit)
    bool gemm_filter = true;              // knob: sample + filter path for tcgen05 batches with top-k beyond register lists
    int scan_shared = -1;                 // knob: SIMT scan selection: 0 = warp-private buffers, otherwise one CTA-shared buffer
    int dense_stage_cap = 0;              // 0 = as many ring stages as fit; set while a sparse CTA must co-reside
    int thr_par = 0;                      // which of the two threshold sets in ws.thr the dense scan uses (pipelined mode)
    bool scan_dynamic = false;            // knob (B200RAG_SCAN_DYNAMIC): dense-scan tiles claimed from a grid-wide counter
                                          // (measured: 28 % SLOWER at 10M rows -- 148 producers on one L2 atomic -- off)
    int scan_ctas = 0;                    // knob (B200RAG_SCAN_CTAS): dense-scan grid size (0 = one CTA per SM)
    int slack = 0;
    int dense_path = 0;  // 0 = auto (SIMT scan for <= 2 queries, tcgen05 GEMM above), 1 = SIMT, 2 = tcgen05
    // Pipelined mode (b200rag_set_pipeline): ONLY the dense scan runs on `stream`; the sparse leg, both legs' tails,
    // exchange and fuse run on `side_stream`, so the next search's scan starts the moment this one's ends.  Candidate
    // lists are double-buffered per call parity and ordered with ev_scan / ev_tail.
    int pipeline = 0;
    bool pipeline_paused = false;         // pipelined mode stays set up, but searches take the classic form on `stream`
    bool legs_classic = false;            // pipelined mode: the last legs took the classic form (heavy tails, tcgen05 batches,
                                          // exhaustive legs): its exchange + fuse stay on `stream`, the hand-over to the
                                          // second stream follows the fuse
    cudaStream_t pipe_stream = nullptr;   // the second stream of pipelined mode: the caller's, or side_stream
    cudaEvent_t ev_scan[2] = {nullptr, nullptr}, ev_tail[2] = {nullptr, nullptr};
    bool ev_tail_rec[2] = {false, false};
    bool exhaustive = false;      // legs score EVERY eligible row canonically and sort (always exact; exact.cu)
    bool exact_fallback = true;   // b200rag_search falls back to the exhaustive pass when the slack guard never clears

    // dense rows
    int64_t n_rows = 0;
    DevBuf dense;  // [n_rows, dim] bf16 bits
    // optional 8-bit copy of the rows for the candidate scan (dense_q8.cu): [n_rows][dim + 16] int8 values + {scale, l1, e2}
    DevBuf dense_q8;
    bool q8 = false;              // b200rag_set_compression
    int64_t q8_rows = 0;          // rows quantised so far (== n_rows whenever q8 is on)
    int q8_slack = 236;           // extra candidates of the 8-bit scan: rows inside its error band, at least 3 L (knob B200RAG_Q8_SLACK)
    bool tail_beside_scan = false; // transient (run_legs -> launch_leg_tail): the tail must co-reside with a dense scan
    bool q8_pipeline = true;      // 8-bit scan in the pipelined form too (knob B200RAG_Q8_PIPELINE=0: classic form)
    DevBuf row_ids;               // i64 [n_rows] global id of every local row, strictly increasing (R1, R5)
    int64_t last_id = INT64_MIN;  // largest id stored so far

    // forward (doc-major) sparse index, kept for the exact re-score and for rebuilds
    int64_t nnz = 0;
    DevBuf fwd_ptr;    // i64 [n_rows+1]
    DevBuf fwd_terms;  // u32 [nnz]
    DevBuf fwd_w;      // f32 [nnz]
    float w_absmax = 0.f;   // max |w| over the first wmax_nnz postings (bounds sparse scores: fixed-point scale of the scan)
    int64_t wmax_nnz = 0;

    // block-major inverted index: block b covers local docs [b*R, (b+1)*R)
    int64_t built_rows = 0;  // rows covered by complete blocks + the trailing partial block as of last build
    int64_t n_blocks = 0;
    int64_t inv_nnz = 0;  // postings stored (block bases padded to 8)
    DevBuf dir;           // u32 [n_blocks, vocab+1]  offsets relative to the block base
    DevBuf blk_base;      // i64 [n_blocks+1]
    DevBuf post_doc;      // u16 [inv_nnz] doc id within block
    DevBuf post_w;        // f32 [inv_nnz]
    std::vector<int64_t> h_blk_base;

    // masks
    std::map<int32_t, DevBuf> masks;
    std::map<int32_t, int64_t> mask_rows;

    // staged query batches; `q`, `h_masks`, `staged_q_terms` and ws.q_* describe the ACTIVE slot
    std::vector<QuerySlot> slots;
    b200rag_query q{};
    bool staged = false;
    std::vector<const uint32_t*> h_masks;
    int64_t staged_q_terms = 0;
    std::vector<int64_t> h_q_indptr;

    Workspace ws;
    b200rag_stats stats{};
    bool profile = false;
    // profiling events: dense scan begin/end, sparse scan begin/end, legs entry, fuse end; `ev` points at the set of the
    // current legs call inside a ring of kProfileRing sets, so back-to-back (unsynchronised) searches keep their timings
    cudaEvent_t ev_ring[kProfileRing][6] = {};
    bool ev_flags[kProfileRing][4] = {};       // dense, sparse, in, out recorded
    cudaEvent_t* ev = ev_ring[0];
    int64_t legs_calls = 0;
    bool ev_dense = false, ev_sparse = false, ev_in = false, ev_out = false;

    // peer-memory exchange window (b200rag_p2p_*): [2 parities][world][slot_bytes] | [world] flags, 128 bytes apart
    void* xwin = nullptr;
    std::vector<void*> xpeers;            // every rank's window as seen from this device (own window at [rank])
    int x_rank = 0, x_world = 0;
    int64_t x_slot_bytes = 0;
    unsigned long long x_epoch = 0;
    long long x_timeout_cycles = 4000000000ll;   // fuse gives up on a peer after this many SM cycles (~2 s)
    cudaStream_t x_stream = nullptr;      // exchange + fuse stream (nullptr = the shard's stream)

    // pinned host staging for results
    void* h_pinned = nullptr;
    size_t h_pinned_cap = 0;
};

// ---- dense_scan.cu -------------------------------------------------------------------------------------
// SIMT bulk-copy scan: approximate fp32 scores, per-CTA top-Lc key lists.  Returns number of lists per query.
int launch_dense_scan(Shard* s, int batch, int Lc, uint64_t* out_lists /*[batch, nlists, Lc]*/, int* nlists);
int dense_scan_nlists(const Shard* s);

// ---- dense_q8.cu ----------------------------------------------------------------------------------------
// 8-bit candidate scan (opt-in): quantise rows [row0, row0 + n) of s->dense into s->dense_q8; scan = launch_dense_scan with
// upper-bound keys (a row outside the retained candidates scores at most the weakest retained key: guard eps = 0)
bool dense_q8_supported(const Shard* s);
int launch_quantize_rows(Shard* s, int64_t row0, int64_t n);
int launch_dense_scan_q8(Shard* s, int batch, int Lc, uint64_t* out_lists, int* nlists);

// ---- dense_umma.cu --------------------------------------------------------------------------------------
// tcgen05 + TMA GEMM with the top-k fused in the TMEM epilogue (<= 128 queries per corpus pass). Same output format.
int launch_dense_gemm(Shard* s, int batch, int Lc, uint64_t* out_lists, int* nlists, float* dbg_scores);
int dense_gemm_nlists(const Shard* s);
// Large top-k with query batches: 1/16 sample pass -> fixed per-query thresholds -> filter pass appending to per-query
// pools -> best Lc of each pool in approx[batch][Lc] (sorted).  Flags `ambiguous` when a pool under- or overflowed.
int launch_dense_gemm_filtered(Shard* s, int batch, int Lc, uint64_t* scratch_a, uint64_t* scratch_b, uint64_t* approx,
                               int32_t* ambiguous);

// ---- select.cu -----------------------------------------------------------------------------------------
// Reduce [batch, n_lists, Lc] key lists to [batch, Lc] (sorted desc) with a tree of smem bitonic merges.
// `a` holds the input; `a`/`b` are used ping-pong; *result points at the final [batch, Lc] list.
int launch_merge_tree(Shard* s, int batch, int n_lists, int Lc, uint64_t* a, uint64_t* b, uint64_t** result);
// q0: first query of the staged batch these `batch` lists belong to; drop_untouched (sparse): a row sharing no term
// with the query gets key 0 (the scans never select such rows; the exhaustive path feeds every row)
int launch_rescore_dense(Shard* s, int batch, int64_t Lc, const uint64_t* approx, uint64_t* exact, int q0 = 0);
int launch_rescore_sparse(Shard* s, int batch, int64_t Lc, const uint64_t* approx, uint64_t* exact, int q0 = 0,
                          bool drop_untouched = false);
// sort exact keys, apply threshold, slack guard, emit b200rag_cand [batch, L]
int launch_finalize_leg(Shard* s, int batch, int Lc, int L, const uint64_t* approx, const uint64_t* exact,
                        float eps_abs, float eps_rel, const float* eps_abs_q /*[batch] or null*/, int has_thr, float thr,
                        b200rag_cand* out, int32_t* ambiguous);
// merge (all levels) + exact re-score + finalize in one launch, when n_lists * Lc keys fit (leg_tail_fits)
bool leg_tail_fits(int n_lists, int Lc);
// gkey: [batch] the scan's final grid-wide threshold keys (0 = none) or nullptr
int launch_leg_tail(Shard* s, bool sparse, int batch, int n_lists, int Lc, int L, const uint64_t* lists, float eps_abs,
                    float eps_rel, const float* eps_abs_q, int has_thr, float thr, b200rag_cand* out, int32_t* ambiguous,
                    const uint64_t* gkey = nullptr);
int launch_fuse(Shard* s, int mode, int batch, int L, int top_k, int rrf_k, const b200rag_cand* gathered,
                int n_shards, int has_trailer, int64_t* out_ids, double* out_scores, int32_t* out_counts,
                int64_t shard_stride_override = 0, const unsigned long long* wait_flags = nullptr,
                unsigned long long wait_epoch = 0);
int launch_exchange(Shard* s, const void* mine, int64_t nbytes, void* const* peer_windows_dev, int world, int rank,
                    int64_t slot_bytes, int parity, unsigned long long epoch);
constexpr int kFlagStrideU64 = 16;   // exchange flags sit 128 bytes apart

// ---- exact.cu ------------------------------------------------------------------------------------------
// Exhaustive exact leg: canonical score of every eligible row, full sort, first L -> out[batch][L].  Never ambiguous.
int launch_exhaustive_leg(Shard* s, bool sparse, int batch, int L, int has_thr, float thr, b200rag_cand* out);
// Drop the rows whose keep bit is clear (keep_words on the device, one bit per local row).
int compact_rows(Shard* s, const uint32_t* keep_words_dev);
int launch_fill_row_ids(Shard* s, int64_t* dst, int64_t first_id, int64_t n);

// ---- sparse.cu -----------------------------------------------------------------------------------------
int build_inverted(Shard* s);
int launch_sparse_scan(Shard* s, int batch, int Lc, uint64_t* out_lists /*[batch, nlists, Lc]*/, float* q_eps /*[batch]*/,
                       int* gthr /*[batch], preset to INT_MIN*/, uint64_t* gkey /*[batch], preset to 0, or null*/);
int sparse_scan_nlists(const Shard* s, int batch, int Lc);

// ---- synth.cu ------------------------------------------------------------------------------------------
int launch_synth_dense(cudaStream_t st, uint64_t seed, int64_t row0, int64_t n, int dim, uint16_t* out);
int launch_synth_sparse(cudaStream_t st, uint64_t seed, int64_t row0, int64_t n, int vocab, int doc_tokens,
                        const uint64_t* thr, const float* idf, const float* tff, int64_t term_mul, int64_t* counts,
                        const int64_t* indptr, uint32_t* terms, float* w);
int launch_exclusive_scan_i64(cudaStream_t st, const int64_t* in, int64_t n, int64_t* out);
int launch_synth_collection_mask(cudaStream_t st, uint64_t seed, int64_t row0, int64_t n, const uint64_t* thr,
                                 int n_coll, int coll, uint32_t* out_words);

}  // namespace b200rag
