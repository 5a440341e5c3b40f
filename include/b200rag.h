/* b200rag.h -- C ABI of the B200-native hybrid retrieval shard engine.
 *
 * The reference (mohammedadnansohail1-pixel/audio-rag) has NO native boundary: its retrieval plugin
 * `QdrantRetriever` (src/audio_rag/retrieval/qdrant.py:14-381) is Python and hands every request to the
 * third-party qdrant-client.  This header is therefore the FFI a maintainer would bind *instead of*
 * `qdrant_client.QdrantClient` inside that plugin; each entry point names the reference call it replaces.
 * The Python binding that does so lives in audio-rag_b200/b200rag/_ffi.py (ctypes) and the plugin mirror
 * in audio-rag_b200/b200rag/retriever.py; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (B200RAG_OK == 0) unless noted;
 *     `b200rag_last_error()` returns a thread-local message for the last non-OK status
 *     (the plugin turns it into audio_rag.core.RetrievalError, qdrant.py:351-352).
 *   - one `b200rag_shard` == one row-range shard of the corpus on one GPU (one process per GPU).
 *     Local row id == insertion order within the shard; global id = cfg.row_base + local.
 *   - "host" pointers are ordinary (ideally pinned) host memory, "dev" pointers are device memory on
 *     the shard's GPU.  The library never takes ownership of caller memory.
 *   - a shard is safe for serialised use from one thread at a time (the reference calls its retriever
 *     from a single blocking handler, api/v1/query.py:18-27,104-115).
 *   - there is no CPU fallback: every compute entry point fails with B200RAG_ERR_NOGPU without a device.
 */
#ifndef B200RAG_H
#define B200RAG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RAG_OK 0
#define B200RAG_ERR_INVALID 1 /* bad argument                                   */
#define B200RAG_ERR_CUDA 2    /* a CUDA call failed (message has the cuda error)  */
#define B200RAG_ERR_NOGPU 3   /* no usable sm_100 device                          */
#define B200RAG_ERR_OOM 4     /* device allocation failed                         */
#define B200RAG_ERR_STATE 5   /* call order violated (e.g. search before build)   */
#define B200RAG_ERR_INEXACT 6 /* the slack guard never cleared and the exhaustive exact pass is disabled:
                                 the result could differ from the exact top-k, so none is returned            */

/* search_type of QdrantRetriever.search (qdrant.py:233,250,272,299,313) */
#define B200RAG_DENSE 0
#define B200RAG_SPARSE 1
#define B200RAG_HYBRID 2

#define B200RAG_MAX_TOPK 256 /* RetrievalConfig.top_k <= 100 (config/schema.py:62); hybrid legs are 2*top_k */

typedef struct b200rag_shard b200rag_shard;

typedef struct {
    int32_t device;         /* CUDA ordinal                                                             */
    int32_t dim;            /* embedding_dim (qdrant.py:24,99): multiple of 256, <= 1024                  */
    int32_t vocab;          /* sparse index space (XLM-R ids from BGE-M3, embeddings/bge.py:95-102)        */
    int32_t docs_per_block; /* doc-range block of the inverted index: power of two, 1024..16384 (0 = 8192) */
    int64_t row_base;       /* global id of local row 0                                                  */
    int64_t reserve_rows;   /* optional pre-allocation hints (0 = grow on demand)                        */
    int64_t reserve_postings;
} b200rag_config;

/* One candidate of one leg as exchanged between shards (16 bytes, all-gathered as raw bytes). */
typedef struct {
    int64_t id;     /* global row id                                 */
    float score;    /* exact leg score (fp32 of the fp64 canonical sum) */
    uint32_t valid; /* 0 = padding                                    */
} b200rag_cand;

/* A batch of queries: the arguments of QdrantRetriever.search (qdrant.py:228-234) for `batch` queries. */
typedef struct {
    int32_t mode;          /* B200RAG_DENSE | SPARSE | HYBRID, after the plugin applied the fallback rules  */
    int32_t batch;         /* number of queries (reference: 1)                                          */
    int32_t top_k;         /* limit (qdrant.py:249,296,311,320)                                          */
    int32_t rrf_k;         /* RRF ranking constant; 0 = qdrant's 2 (hybrid/fusion.py)                     */
    int32_t has_threshold; /* score_threshold, dense legacy collections only (qdrant.py:331)            */
    float score_threshold;
    const uint16_t* q_dense_bits; /* host [batch, dim] bf16 bits of the UNIT query (b200rag_normalize_bf16)   */
    const int64_t* q_sp_indptr;   /* host [batch+1]  (NULL when mode == DENSE)                               */
    const uint32_t* q_sp_terms;   /* host, ascending & unique per query                                    */
    const float* q_sp_weights;    /* host                                                                  */
    const int32_t* mask_ids;      /* host [batch] eligibility mask per query, -1 = all rows; NULL = none   */
} b200rag_query;

/* ---- library ---------------------------------------------------------------------------------------- */
const char* b200rag_version(void);
const char* b200rag_last_error(void);
int b200rag_device_count(void); /* number of sm_100 devices visible (0 without a GPU; never fails) */

/* Host-only helper (no GPU): cosine pre-normalisation + bf16 rounding, the arithmetic of SURVEY R2.
 * Replaces the normalisation qdrant applies to COSINE vectors (collection schema at qdrant.py:98-117).
 * ss = sequential fp64 sum of squares; y = fp32(fp64(x)/sqrt(ss)); bits = RNE bf16(y); zero rows pass through. */
int b200rag_normalize_bf16(const float* x_host, int64_t n, int32_t dim, uint16_t* out_bits_host);

/* ---- shard lifetime  (replaces QdrantClient(...) construction, qdrant.py:35-54) ------------------------ */
int b200rag_shard_create(const b200rag_config* cfg, b200rag_shard** out);
void b200rag_shard_destroy(b200rag_shard* s);
/* Run on the caller's cudaStream_t (NULL == the legacy default stream).  Until this is called the shard
 * uses a private non-blocking stream. */
int b200rag_set_stream(b200rag_shard* s, void* cuda_stream);
int b200rag_set_slack(b200rag_shard* s, int32_t slack);      /* extra approximate candidates per leg before the
                                                                exact re-score (0 = default max(16, L/2))      */
/* Exactness contract.  The scans select Lc = L + slack candidates by APPROXIMATE score and re-score them exactly; a
 * guard flags a leg as ambiguous when a row outside the candidates could still belong to the exact top-L (massive
 * ties, near-duplicate scores around the cut).  b200rag_search then widens the slack and repeats; if the guard still
 * has not cleared when the slack reaches its cap, the legs are recomputed EXHAUSTIVELY: every eligible row is scored in
 * the canonical order and the top-L are taken from a full sort (always exact, ~10x the cost of a scan).
 *   set_exhaustive(1)      make b200rag_legs take the exhaustive path directly (multi-shard callers use it for their own
 *                          last retry; tests use it as an independent in-library cross-check of the scan kernels);
 *   set_exact_fallback(0)  disable the automatic exhaustive pass: b200rag_search then fails with
 *                          B200RAG_ERR_INEXACT instead of returning a result whose guard never cleared
 *                          (default 1; the environment variable B200RAG_EXACT_FALLBACK=0 sets the default to 0). */
int b200rag_set_exhaustive(b200rag_shard* s, int32_t on);
/* Corpus compression for the candidate scan (opt-in, dim 512 or 1024).  The shard keeps an 8-bit copy
 * of its rows ([n][dim + 16]: symmetric int8 per row + {scale, l1, residual norm}) beside the bf16 rows; searches of 1-2
 * queries scan THAT (half the bytes), ranking rows by a rigorous UPPER BOUND of their exact score, keep a wider candidate
 * set (the rows inside the quantisation error band: max(`B200RAG_Q8_SLACK` = 236, 3 L) extra candidates) and re-score the
 * candidates exactly from the bf16 rows -- ids and scores stay bit-identical to the uncompressed path; when the guard does
 * not clear, the retry scans the bf16 rows.  Costs (dim + 16) bytes per row of HBM.  Maintained through add / compact /
 * load.  No reference counterpart (the reference's collections are created without quantisation, qdrant.py:95-118;
 * Qdrant's own scalar quantisation with rescoring is approximate, this is not). */
int b200rag_set_compression(b200rag_shard* s, int32_t on);
int b200rag_set_exact_fallback(b200rag_shard* s, int32_t on);
/* Pipelined searches (throughput mode for callers that enqueue search after search on staged batches: b200rag_legs ->
 * [b200rag_p2p_exchange ->] b200rag_fuse / b200rag_p2p_fuse, no host synchronisation in between).  With pipeline on, ONLY
 * the dense scan of a search runs on the shard's stream; the sparse leg, both legs' merge + exact re-score + finalize,
 * the candidate exchange and the fuse run on an internal second stream, so the next search's scan starts the moment
 * this one's ends (candidate lists are double-buffered, ordering is by CUDA events inside the library).  Results of a
 * search are complete on b200rag_result_stream(): enqueue or synchronise dependent work THERE (everything that writes
 * or reads the caller's candidate / output buffers runs on that one in-order stream, so single buffers suffice).
 * `second_stream`: a cudaStream_t of the caller to use as that second stream (it must outlive the pipelined mode), or
 * NULL for a stream the library owns.  b200rag_search (synchronous, host buffers) always runs the classic form. */
int b200rag_set_pipeline(b200rag_shard* s, int32_t on, void* second_stream);
void* b200rag_result_stream(const b200rag_shard* s); /* cudaStream_t on which fused results become available */
/* While paused, a pipelined shard runs its searches in the classic form, everything on the shard's stream (the first
 * such search waits, on the device, for the pipelined ones still in flight).  No CUDA call, no synchronisation: a caller
 * that mixes a stream of pipelined searches with an occasional synchronous one (ShardedSearcher.search) brackets the
 * latter with pause(1) / pause(0) -- a lone search is ~10-50 us faster without the stream hand-overs. */
int b200rag_pipeline_pause(b200rag_shard* s, int32_t on);
/* Dense kernel choice: 0 = auto (bulk-copy SIMT scan for <= 2 queries, tcgen05 GEMM above), 1 = SIMT scan,
 * 2 = tcgen05 GEMM.  Both produce bit-identical results (candidates are re-scored in the canonical order). */
int b200rag_set_dense_path(b200rag_shard* s, int32_t path);
/* Test hook: run the tcgen05 GEMM on the staged batch and write its raw approximate scores [batch, rows] (fp32). */
int b200rag_debug_dense_scores(b200rag_shard* s, float* out_scores_dev);
int b200rag_sync(b200rag_shard* s);

/* ---- ingest  (replaces client.upsert, qdrant.py:197-220) ---------------------------------------------- */
/* Append n rows.  dense_bits: [n, dim] unit rows as bf16 bits.  Sparse part is doc-major CSR
 * (indptr[n+1] from 0, terms ascending & unique per row, all < vocab); pass NULLs for dense-only rows. */
int b200rag_add(b200rag_shard* s, int64_t n, const uint16_t* dense_bits_host, const int64_t* sp_indptr_host,
                const uint32_t* sp_terms_host, const float* sp_weights_host);
/* Same with device-resident inputs (bulk / synthetic ingest). nnz = indptr[n]. */
int b200rag_add_device(b200rag_shard* s, int64_t n, const uint16_t* dense_bits_dev, const int64_t* sp_indptr_dev,
                       const uint32_t* sp_terms_dev, const float* sp_weights_dev, int64_t nnz);
/* The same two calls with explicit GLOBAL row ids (host array [n]; NULL = cfg.row_base + local row, which is what
 * b200rag_add / b200rag_add_device assign).  Ids must be strictly increasing over the shard's lifetime, so that the
 * order of local rows is the order of global ids and the tie-break of SURVEY R5 (smaller row id first) is the same
 * on one shard and on any row-partition of the corpus.  A plugin that spreads one insertion-ordered row space over
 * several shards (B200Retriever with shards > 1) routes each add() batch to a shard and passes the rows' global ids. */
int b200rag_add_ids(b200rag_shard* s, int64_t n, const uint16_t* dense_bits_host, const int64_t* sp_indptr_host,
                    const uint32_t* sp_terms_host, const float* sp_weights_host, const int64_t* ids_host);
int b200rag_add_device_ids(b200rag_shard* s, int64_t n, const uint16_t* dense_bits_dev, const int64_t* sp_indptr_dev,
                           const uint32_t* sp_terms_dev, const float* sp_weights_dev, int64_t nnz,
                           const int64_t* ids_host);
/* GPU-side ingest pack: RAW fp32 vectors from the host (what an embedder returns, embeddings/bge.py:125)
 * are copied to the device and normalised + rounded to bf16 there (b200rag_normalize_bf16_device: bit-equal to the
 * host routine), so add() does not spend host time on the corpus rows.  CSR and ids as in b200rag_add_ids. */
int b200rag_add_f32(b200rag_shard* s, int64_t n, const float* dense_f32_host, const int64_t* sp_indptr_host,
                    const uint32_t* sp_terms_host, const float* sp_weights_host, const int64_t* ids_host);
/* Physically drop rows (delete_collection, qdrant.py:354-363: the reference drops the collection's storage).
 * keep_words_host: bit (r & 31) of word (r >> 5) set <=> local row r survives; n_rows_mask must equal the shard's row
 * count.  Surviving rows keep their order and their global ids; dense rows, the forward index and the id map are
 * rewritten on the device, the inverted index is rebuilt on the next sparse search (or b200rag_build).  Every mask
 * and staged batch of the shard is dropped (their bit positions are local rows). */
int b200rag_compact(b200rag_shard* s, const uint32_t* keep_words_host, int64_t n_rows_mask);
/* Bring the block-major inverted index up to date with the rows added so far (incremental). */
int b200rag_build(b200rag_shard* s);
int64_t b200rag_count(const b200rag_shard* s);    /* rows in the shard (client.get_collection().points_count, qdrant.py:369-370) */
int64_t b200rag_postings(const b200rag_shard* s); /* sparse non-zeros in the shard */
/* Drop every row (delete_collection of the last collection, qdrant.py:354-363); keeps allocations. */
int b200rag_clear(b200rag_shard* s);
/* Test/debug read-back of stored rows (the oracle checks of tests/ and bench.py score exactly these bits).
 * read_sparse: indptr_out_host[n+1] (from 0) is always written; terms / weights only when both are non-NULL and
 * cap_nnz >= indptr_out_host[n] (else B200RAG_ERR_INVALID, so a caller can size its arrays from a first call). */
int b200rag_read_dense(b200rag_shard* s, int64_t row, int64_t n, uint16_t* out_bits_host);
int b200rag_read_sparse(b200rag_shard* s, int64_t row, int64_t n, int64_t* indptr_out_host, uint32_t* terms_out_host,
                        float* weights_out_host, int64_t cap_nnz);
int b200rag_read_row_ids(b200rag_shard* s, int64_t row, int64_t n, int64_t* ids_out_host);

/* ---- eligibility masks  (replace Filter(must=[FieldCondition...]) + the collection itself, qdrant.py:262-269) */
/* Bit (r & 31) of word (r >> 5) set <=> local row r is eligible.  n_rows bits are read; rows beyond are ineligible. */
int b200rag_mask_set(b200rag_shard* s, int32_t mask_id, const uint32_t* words_host, int64_t n_rows);
int b200rag_mask_set_device(b200rag_shard* s, int32_t mask_id, const uint32_t* words_dev, int64_t n_rows);
int b200rag_mask_drop(b200rag_shard* s, int32_t mask_id);

/* ---- search  (replaces client.query_points, qdrant.py:281-332) ------------------------------------------ */
/* Whole path on one shard with HOST buffers: stage -> legs -> fuse -> read back.
 * out_ids [batch, top_k] global row ids, out_scores [batch, top_k] (fp64: leg score, or RRF fused score),
 * out_counts [batch] results per query.  Returned order: score desc, stated tie-breaks. */
int b200rag_search(b200rag_shard* s, const b200rag_query* q, int64_t* out_ids_host, double* out_scores_host,
                   int32_t* out_counts_host);

/* The same path in its three device stages, for multi-shard composition (one process per GPU):
 *   stage : copy the query batch to the device (async on the shard's stream);
 *   legs  : per-shard candidates, exact-scored & ordered; cands_dev is [nlegs, batch, L] b200rag_cand with
 *           nlegs = 2 (dense, sparse) for HYBRID else 1, L = 2*top_k for HYBRID else top_k;
 *           ambiguous_dev (device int32, may be NULL) is zeroed, then incremented when a leg's slack guard fails;
 *   fuse  : merge `n_shards` gathered candidate sets [n_shards, nlegs, batch, L] under R5, then RRF under
 *           R9/R10 (HYBRID), writing device results [batch, top_k].
 * Exchange format: with has_trailer != 0 every shard's block is followed by ONE trailer b200rag_cand whose
 * `id` low word is that shard's ambiguity counter (pass &trailer as ambiguous_dev to b200rag_legs), i.e. the
 * gathered buffer is [n_shards, nlegs*batch*L + 1]; fuse then also writes the sum of the counters to
 * out_counts_dev[batch] (so out_counts_dev holds batch + 1 ints). */
int b200rag_stage(b200rag_shard* s, const b200rag_query* q);
/* Several staged batches at once (slot 0 is the one b200rag_stage uses): stage_slot copies a batch into its own device
 * block and makes it the active one; use_slot re-activates an already staged batch without any copy, so a caller can
 * keep a queue of query batches resident in HBM and enqueue legs/exchange/fuse for them back to back with no host
 * synchronisation in between.  Staged batches are invalidated by b200rag_mask_set / b200rag_mask_drop / b200rag_clear. */
int b200rag_stage_slot(b200rag_shard* s, const b200rag_query* q, int32_t slot);
int b200rag_use_slot(b200rag_shard* s, int32_t slot);
/* Device-resident queries (the embedder's outputs never leave the GPU; embeddings/bge.py:137-157 moves them
 * to Python lists today).  normalize_bf16_device is the device twin of b200rag_normalize_bf16 (bit-equal output: the
 * same fp64 operations in the same order, one thread per row).  stage_device is b200rag_stage_slot with q_dense_bits,
 * q_sp_terms and q_sp_weights pointing to DEVICE memory (device-to-device copies on the shard's stream);
 * q_sp_indptr and mask_ids stay host arrays (batch + 1 small integers the caller knows anyway).  The sparse terms are
 * validated on the device (in range, ascending and unique per query). */
int b200rag_normalize_bf16_device(b200rag_shard* s, const float* x_dev, int64_t n, uint16_t* out_bits_dev);
int b200rag_stage_device(b200rag_shard* s, const b200rag_query* q, int32_t slot);
int b200rag_legs_len(const b200rag_query* q, int32_t* nlegs, int32_t* L);
int b200rag_legs(b200rag_shard* s, void* cands_dev, int32_t* ambiguous_dev);
int b200rag_fuse(b200rag_shard* s, const void* gathered_dev, int32_t n_shards, int32_t has_trailer,
                 int64_t* out_ids_dev, double* out_scores_dev, int32_t* out_counts_dev);

/* ---- persistence (the reference relies on Qdrant's volume, docker-compose.yml:36-37) -------------------------------
 * One file per shard: header | dense bf16 rows | forward sparse index (indptr, terms, weights) | global row ids -- raw,
 * contiguous arrays at offsets the 40-byte header determines, i.e. mmap-able; b200rag/shardfile.py documents the layout
 * byte by byte and reads (numpy.memmap) and writes it without a GPU.  The inverted index,
 * directories and the weight bound are rebuilt on load (~0.2 s per 10M rows) so the file has no layout the kernels
 * depend on.  `load` needs an EMPTY shard created with the same dim and vocab; masks and payloads are the plugin's
 * (B200Retriever.save/load write them next to this file).  Returns B200RAG_ERR_INVALID on a foreign, mismatching
 * (dim / vocab) or truncated file; the CONTENT of a file that passes those checks is trusted like the arguments of
 * b200rag_add are (indptr monotone, terms ascending and < vocab) -- it is this library's own output. */
int b200rag_save(b200rag_shard* s, const char* path);
int b200rag_load(b200rag_shard* s, const char* path);

/* ---- peer-memory candidate exchange (one process per GPU, NVLink/NVSwitch P2P) ---------------------------------
 * Replaces the NCCL all-gather between `legs` and `fuse` when every shard of a search sits on a GPU of the same box:
 *   export  : allocate this shard's exchange window (2 parities x world slots of slot_bytes + flags) and return its
 *             CUDA IPC handle (B200RAG_IPC_HANDLE_BYTES bytes) for the other ranks;
 *   attach  : open the `world` handles (rank-ordered, own handle included) -> peer pointers;
 *   exchange: ONE kernel stores this rank's candidate block (nbytes <= slot_bytes, trailer included) into slot `rank`
 *             of EVERY rank's window with peer stores, then publishes the search's epoch to every rank's flag word
 *             (release at system scope);
 *   fuse    : b200rag_fuse on the local window; the kernel itself waits (acquire at system scope) until all `world`
 *             flags carry the epoch, so there is no host synchronisation and no collective launch in the search.
 * Every rank must make the same sequence of exchange/fuse calls (SPMD).  A rank may run at most one search ahead
 * of its peers (its next fuse waits for their next epoch), which the two slot parities cover.
 * p2p_fuse writes batch + 2 ints to out_counts_dev: [batch] = sum of the shards' ambiguity counters (or -1), and
 * [batch + 1] = a STICKY timeout latch which the caller zeroes once when it allocates the buffer: if a peer's flag
 * does not arrive within the timeout (B200RAG_P2P_TIMEOUT_MS, default 2000) every block that gave up sets the latch
 * and empties its query, so a partial timeout can never pass for a result.  After a timeout the ranks' epochs have
 * diverged: close and re-export / re-attach the windows (ShardedSearcher.reset) before searching again. */
#define B200RAG_IPC_HANDLE_BYTES 64
int b200rag_p2p_export(b200rag_shard* s, int32_t world, int64_t slot_bytes, uint8_t* handle_out);
int b200rag_p2p_attach(b200rag_shard* s, int32_t rank, int32_t world, const uint8_t* handles);
int b200rag_p2p_exchange(b200rag_shard* s, const void* mine_dev, int64_t nbytes);
int b200rag_p2p_fuse(b200rag_shard* s, int64_t* out_ids_dev, double* out_scores_dev, int32_t* out_counts_dev);
int b200rag_p2p_close(b200rag_shard* s);
/* Optional: launch exchange + fuse on `cuda_stream` instead of the shard's stream, so that a caller can let the NEXT
 * search's scans start while this search still waits for its peers (the caller orders the two streams with events:
 * exchange after this search's legs; buffers reused only after the fuse that read them).  NULL = the shard's stream. */
int b200rag_p2p_set_stream(b200rag_shard* s, void* cuda_stream);

/* ---- shard group: several shards of one corpus driven by ONE process ----------------------------------------------
 * The reference's retriever is one in-process object (pipeline/orchestrator.py:48-74), so behind the plugin the row
 * sharding of the corpus over the GPUs of a box lives in the caller's process: a group holds `n` shards (one per GPU,
 * or several per GPU), and group_search is b200rag_search over all of them -- on every shard stage + legs on a
 * per-device worker thread, the candidate blocks gathered into the first shard's window by asynchronous peer copies
 * ordered with CUDA events, then fuse + read-back on the first shard.  The result is bit-identical to one shard holding
 * all rows (global ids and R5 order come from the shards' row ids, see b200rag_add_ids).  The shards stay owned by the
 * caller (add / mask_set / compact are per shard; q->mask_ids name the SAME mask id on every shard) and must outlive
 * the group.  Same retry / exhaustive-fallback / B200RAG_ERR_INEXACT contract as b200rag_search. */
typedef struct b200rag_group b200rag_group;
int b200rag_group_create(b200rag_shard* const* shards, int32_t n, b200rag_group** out);
void b200rag_group_destroy(b200rag_group* g);
int32_t b200rag_group_size(const b200rag_group* g);
int b200rag_group_search(b200rag_group* g, const b200rag_query* q, int64_t* out_ids_host, double* out_scores_host,
                         int32_t* out_counts_host);

/* Counters of the last `legs` call, for bench.py's gpu_launches / roofline bookkeeping. */
typedef struct {
    int32_t kernel_launches;     /* kernels of this library launched by the last legs+fuse               */
    int32_t dense_path;          /* 0 = none, 1 = SIMT bulk-copy scan, 2 = tcgen05 GEMM, 3 = 8-bit SIMT scan */
    int64_t dense_bytes;         /* algorithmic bytes of the dense leg (rows * dim * 2 per corpus pass; rows * (dim + 16) for the 8-bit scan) */
    int64_t sparse_postings;     /* postings of the query terms in this shard (sum over batch)            */
    int32_t dense_passes;        /* corpus passes made for the batch                                      */
    int32_t retries;             /* slack-guard retries inside b200rag_search                              */
    float dense_scan_ms;         /* with profiling on: device time of the dense scan kernel(s) of the last legs */
    float sparse_scan_ms;        /* ... and of the sparse scan kernel (CUDA events on the shard's stream)      */
    float pre_scan_ms;           /* ... from the entry of `legs` to the start of the dense scan (launch latency)  */
    float tail_ms;               /* ... from the end of the dense scan to the end of the last `fuse`               */
    int32_t exhaustive;          /* 1 = the last legs took the exhaustive exact path (set_exhaustive / fallback)   */
    int32_t reserved;
} b200rag_stats;
int b200rag_get_stats(const b200rag_shard* s, b200rag_stats* out);
int b200rag_group_get_stats(const b200rag_group* g, b200rag_stats* out); /* last group_search: launches summed over shards */
/* With profiling on: event timings of the legs call made `steps_back` calls ago (0 = the last one, < 64), for callers
 * that enqueue many searches before synchronising; only the four *_ms fields of `out` are filled. */
int b200rag_get_stats_step(const b200rag_shard* s, int32_t steps_back, b200rag_stats* out);
/* Bracket the two scan kernels with CUDA events on the launching stream (bench.py's roofline figures). */
int b200rag_set_profiling(b200rag_shard* s, int32_t on);

/* ---- synthetic corpus generation on the device (bench / tests; twins of b200rag/synth.py) -------------- */
int b200rag_synth_dense(b200rag_shard* s, uint64_t seed, int64_t global_row_start, int64_t n, uint16_t* out_bits_dev);
/* Two-pass doc-major CSR generation: call with terms_dev == NULL to get counts_dev[n] (int64), scan them into
 * indptr_dev[n+1] yourself (or use b200rag_exclusive_scan_i64), then call again to fill terms/weights. */
int b200rag_synth_sparse(b200rag_shard* s, uint64_t seed, int64_t global_row_start, int64_t n, int32_t doc_tokens,
                         const uint64_t* zipf_thresholds_dev, const float* idf_dev, const float* tff_dev,
                         int64_t term_mul, int64_t* counts_dev, const int64_t* indptr_dev, uint32_t* terms_dev,
                         float* weights_dev);
int b200rag_exclusive_scan_i64(b200rag_shard* s, const int64_t* in_dev, int64_t n, int64_t* out_dev /* n+1 */);
/* Mask for "collection id == c" over rows [0, n): rows drawn Zipf over n_collections (synth.row_collections). */
int b200rag_synth_collection_mask(b200rag_shard* s, uint64_t seed, int64_t global_row_start, int64_t n,
                                  const uint64_t* coll_thresholds_dev, int32_t n_collections, int32_t collection,
                                  uint32_t* out_words_dev);

#ifdef __cplusplus
}
#endif
#endif /* B200RAG_H */
