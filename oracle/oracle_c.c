/* C restatement of the canonical oracle (oracle/oracle.py) for sizes numpy is too slow for.
 * TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
 * PARITY UNPINNED: see the header of oracle/oracle.py (the arithmetic belongs to the
 * third-party qdrant-client, absent here; the reference has no golden vectors for this path).
 *
 * Follows, rule by rule:
 *   R2  dense score  = fp32( fp64 sum of c_k*q_k in the lane-blocked order of oracle.py )  (qdrant local/distances.py::cosine_similarity
 *                                                                         on the unit vectors the engine stores as bf16)
 *   R3  sparse score = fp32( sum over common indices ascending of fp64(w_q)*fp64(w_d) )
 *                                                                        (local/sparse_distances.py::sparse_dot_product)
 *   R7  a document with no common index is "untouched"
 * Call sites in the reference: src/audio_rag/retrieval/qdrant.py:281-332.
 *
 * Also holds C twins of the synthetic generators (audio-rag_b200/b200rag/synth.py) so that
 * million-row fixtures can be produced on the CPU in seconds.
 *
 * Build: make -C oracle     (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline double bf16_to_f64(uint16_t b) {
    uint32_t u = ((uint32_t)b) << 16;
    float f;
    memcpy(&f, &u, 4);
    return (double)f;
}

void oracle_dense_scores(const uint16_t* bits, int64_t n, int32_t dim, const uint16_t* q, float* out) {
    /* canonical order of oracle.py::dense_scores: 32 lane partials (lane = (k % 256) / 8, ascending k),
       then a pairwise tree over the lanes */
    double qd[4096];
    const int nch = dim / 256;
    for (int k = 0; k < dim; ++k) qd[k] = bf16_to_f64(q[k]);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        const uint16_t* row = bits + r * (int64_t)dim;
        double p[32];
        for (int l = 0; l < 32; ++l) {
            double acc = 0.0;
            for (int c = 0; c < nch; ++c)
                for (int e = 0; e < 8; ++e) {
                    const int k = c * 256 + l * 8 + e;
                    acc += bf16_to_f64(row[k]) * qd[k];
                }
            p[l] = acc;
        }
        for (int w = 32; w > 1; w >>= 1)
            for (int i = 0; i < w / 2; ++i) p[i] = p[2 * i] + p[2 * i + 1];
        out[r] = (float)p[0] + 0.0f;
    }
}

/* q_idx ascending & unique; per-document terms ascending & unique. */
void oracle_sparse_scores(const int64_t* indptr, const uint32_t* terms, const float* w, int64_t n,
                          const int64_t* q_idx, const float* q_val, int32_t nq,
                          float* out, uint8_t* touched) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t d = 0; d < n; ++d) {
        int64_t i = indptr[d], e = indptr[d + 1];
        int j = 0;
        double acc = 0.0;
        uint8_t t = 0;
        while (i < e && j < nq) {
            int64_t a = (int64_t)terms[i], b = q_idx[j];
            if (a == b) { acc += (double)q_val[j] * (double)w[i]; t = 1; ++i; ++j; }
            else if (a < b) ++i;
            else ++j;
        }
        out[d] = (float)acc + 0.0f;
        touched[d] = t;
    }
}

/* ---------------------------------------------------------------- synthetic generator twins */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t stream_key(uint64_t seed, uint64_t stream) { return mix64(seed * 0x10000ull + stream); }
static inline uint64_t row_key(uint64_t skey, uint64_t row) { return mix64(skey ^ (row * 0xD6E8FEB86659FD93ull)); }
static inline int64_t raw_int(uint64_t rkey, uint64_t j) {
    uint64_t h = mix64(rkey + j);
    return (int64_t)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + ((h >> 48) & 0xFFFF)) - 131070;
}
static inline uint16_t f32_to_bf16(float y) {
    uint32_t u;
    memcpy(&u, &y, 4);
    u = u + 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

void oracle_synth_dense_bf16(uint64_t seed, int64_t row_start, int64_t n, int32_t dim, uint16_t* out) {
    uint64_t sk = stream_key(seed, 1);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        uint64_t rk = row_key(sk, (uint64_t)(row_start + r));
        int64_t ss = 0;
        for (int k = 0; k < dim; ++k) { int64_t x = raw_int(rk, (uint64_t)k); ss += x * x; }
        if (ss == 0) ss = 1;
        double nrm = sqrt((double)ss);
        for (int k = 0; k < dim; ++k) {
            float y = (float)((double)raw_int(rk, (uint64_t)k) / nrm);
            out[r * (int64_t)dim + k] = f32_to_bf16(y);
        }
    }
}

static inline int64_t zipf_rank(const uint64_t* thr, int64_t v, uint64_t u) {
    /* number of thresholds <= u  (np.searchsorted side='right'), clipped to v-1 */
    int64_t lo = 0, hi = v;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (thr[mid] <= u) lo = mid + 1; else hi = mid; }
    return lo < v ? lo : v - 1;
}

static void sort_i64(int64_t* a, int n) { /* insertion sort, n <= 256 */
    for (int i = 1; i < n; ++i) { int64_t x = a[i]; int j = i - 1; while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; } a[j + 1] = x; }
}

/* pass 1 (counts != NULL): counts[d] = number of distinct terms; pass 2: fill terms/weights at indptr[d]. */
void oracle_synth_sparse(uint64_t seed, int64_t row_start, int64_t n, int32_t vocab, int32_t doc_tokens,
                         const uint64_t* thr, const float* idf, const float* tff, int64_t term_mul,
                         int64_t* counts, const int64_t* indptr, uint32_t* terms, float* w) {
    uint64_t sk = stream_key(seed, 3);
#pragma omp parallel for schedule(static)
    for (int64_t d = 0; d < n; ++d) {
        int64_t t[256];
        uint64_t rk = row_key(sk, (uint64_t)(row_start + d));
        for (int i = 0; i < doc_tokens; ++i)
            t[i] = (zipf_rank(thr, vocab, mix64(rk + (uint64_t)i) >> 11) * term_mul) % vocab;
        sort_i64(t, doc_tokens);
        int64_t c = 0, o = indptr ? indptr[d] : 0;
        for (int i = 0; i < doc_tokens;) {
            int j = i;
            while (j < doc_tokens && t[j] == t[i]) ++j;
            if (!counts) { terms[o + c] = (uint32_t)t[i]; w[o + c] = idf[t[i]] * tff[j - i]; }
            ++c;
            i = j;
        }
        if (counts) counts[d] = c;
    }
}
