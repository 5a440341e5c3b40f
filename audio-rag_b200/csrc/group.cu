// b200rag_group: several shards of ONE corpus driven from ONE host process (one GPU each, or several on one GPU).
//
// This is the multi-GPU form of the plugin boundary: the reference's retriever is a single in-process object
// (src/audio_rag/pipeline/orchestrator.py:48-74 builds one and hands it to both pipelines), so behind
// B200Retriever.search the 8-way row sharding of SURVEY 8e has to live in the caller's process -- the SPMD
// (one process per GPU) form of the same search is b200rag/dist.py.
//
// One search = on every shard [stage -> legs -> copy of the candidate block into the root shard's gather window ->
// event], then on the root shard [wait for the events -> fuse -> read back].  The per-shard host work (a pinned-memory
// copy of the query, ~5 kernel launches) runs on one persistent worker thread per DEVICE so that 8 GPUs start their
// scans within a few microseconds of each other instead of ~30 us apart; shards that share a device are driven by the
// same worker, in order.  There is no collective and no spinning kernel: the gather is `n` asynchronous peer copies
// ordered by CUDA events, which is also correct when several shards share one GPU (tests on a 1-GPU box).
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "engine.h"

namespace b200rag {

static int default_slack_g(int L) { return L / 2 > 16 ? L / 2 : 16; }

struct Group;

struct GroupWorker {
    Group* g = nullptr;
    int device = 0;
    std::vector<int> shards;                 // indices into Group::shards, in order
    std::thread th;
    std::atomic<uint64_t> go{0}, done{0};
    int rc = B200RAG_OK;
    std::string err;
};

struct Group {
    std::vector<Shard*> shards;
    std::vector<cudaEvent_t> ev;             // per shard: candidate block copied into the root window
    std::vector<GroupWorker*> workers;       // workers[0] is run by the calling thread (the root shard's device)
    std::mutex mu;
    std::condition_variable cv;
    bool quit = false;
    uint64_t gen = 0;
    // current job
    const b200rag_query* q = nullptr;
    size_t block_bytes = 0;                  // candidate block of one shard incl. trailer
    // root side
    DevBuf window;                           // [n][slot] gathered candidate blocks, on the root device
    size_t slot_bytes = 0;
    DevBuf out;                              // ids | scores | counts(+ambiguity)
    void* h_out = nullptr;
    size_t h_out_cap = 0;
    b200rag_stats last_stats{};
};

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}

// stage + legs + copy + event for every shard of one worker
static int run_worker_job(Group* g, GroupWorker* w) {
    const b200rag_query* q = g->q;
    Shard* root = g->shards[0];
    for (int idx : w->shards) {
        Shard* s = g->shards[(size_t)idx];
        B2_TRY(b200rag_stage((b200rag_shard*)s, q));
        B2_TRY(s->ws.cands.ensure(g->block_bytes, 0, s->stream));
        uint8_t* blk = s->ws.cands.as<uint8_t>();
        int32_t* amb = reinterpret_cast<int32_t*>(blk + g->block_bytes - sizeof(b200rag_cand));
        B2_CUDA(cudaMemsetAsync(amb, 0, sizeof(b200rag_cand), s->stream));
        B2_TRY(b200rag_legs((b200rag_shard*)s, blk, amb));
        uint8_t* dst = g->window.as<uint8_t>() + (size_t)idx * g->slot_bytes;
        if (s->cfg.device == root->cfg.device)
            B2_CUDA(cudaMemcpyAsync(dst, blk, g->block_bytes, cudaMemcpyDeviceToDevice, s->stream));
        else
            B2_CUDA(cudaMemcpyPeerAsync(dst, root->cfg.device, blk, s->cfg.device, g->block_bytes, s->stream));
        B2_CUDA(cudaEventRecord(g->ev[(size_t)idx], s->stream));
    }
    return B200RAG_OK;
}

static void worker_main(GroupWorker* w) {
    Group* g = w->g;
    cudaSetDevice(w->device);
    uint64_t seen = 0;
    for (;;) {
        int spins = 0;
        uint64_t gen;
        while ((gen = w->go.load(std::memory_order_acquire)) == seen) {
            if (++spins < 20000) { cpu_relax(); continue; }      // ~1 ms of polling, then sleep
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv.wait_for(lk, std::chrono::milliseconds(50), [&] { return w->go.load(std::memory_order_acquire) != seen; });
        }
        seen = gen;
        if (g->quit) return;
        w->rc = run_worker_job(g, w);
        if (w->rc != B200RAG_OK) w->err = b200rag_last_error();
        w->done.store(gen, std::memory_order_release);
    }
}

// one pass over all shards: legs everywhere, gather, fuse on the root, results in g->h_out
static int group_pass(Group* g, int B, int K, int L, size_t out_bytes, int32_t* ambiguous) {
    ++g->gen;
    for (size_t i = 1; i < g->workers.size(); ++i) g->workers[i]->go.store(g->gen, std::memory_order_release);
    if (g->workers.size() > 1) g->cv.notify_all();
    int rc = run_worker_job(g, g->workers[0]);
    std::string err = rc != B200RAG_OK ? std::string(b200rag_last_error()) : std::string();
    for (size_t i = 1; i < g->workers.size(); ++i) {
        GroupWorker* w = g->workers[i];
        while (w->done.load(std::memory_order_acquire) != g->gen) cpu_relax();
        if (w->rc != B200RAG_OK && rc == B200RAG_OK) { rc = w->rc; err = w->err; }
    }
    if (rc != B200RAG_OK) { set_error(err); return rc; }
    Shard* root = g->shards[0];
    B2_CUDA(cudaSetDevice(root->cfg.device));
    cudaStream_t st = root->stream;
    for (size_t i = 1; i < g->shards.size(); ++i) B2_CUDA(cudaStreamWaitEvent(st, g->ev[i], 0));
    uint8_t* d = g->out.as<uint8_t>();
    const size_t o_sc = (size_t)B * K * 8, o_cnt = 2 * o_sc;
    B2_TRY(launch_fuse(root, root->q.mode, B, L, K, root->q.rrf_k, g->window.as<b200rag_cand>(), (int)g->shards.size(), 1,
                       (int64_t*)d, (double*)(d + o_sc), (int32_t*)(d + o_cnt),
                       (int64_t)(g->slot_bytes / sizeof(b200rag_cand))));
    B2_CUDA(cudaMemcpyAsync(g->h_out, d, out_bytes, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    *ambiguous = ((const int32_t*)((const uint8_t*)g->h_out + o_cnt))[B];
    return B200RAG_OK;
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

int b200rag_group_create(b200rag_shard* const* shards, int32_t n, b200rag_group** out) {
    if (shards == nullptr || out == nullptr || n < 1 || n > 64) { set_error("group_create: bad argument"); return B200RAG_ERR_INVALID; }
    *out = nullptr;
    for (int i = 0; i < n; ++i)
        if (shards[i] == nullptr) { set_error("group_create: null shard"); return B200RAG_ERR_INVALID; }
    Group* g = new Group();
    for (int i = 0; i < n; ++i) g->shards.push_back((Shard*)shards[i]);
    const int root_dev = g->shards[0]->cfg.device;
    for (int i = 1; i < n; ++i) {
        if (g->shards[(size_t)i]->dim != g->shards[0]->dim || g->shards[(size_t)i]->vocab != g->shards[0]->vocab) {
            delete g;
            set_error("group_create: shards differ in dim / vocab");
            return B200RAG_ERR_INVALID;
        }
    }
    g->ev.assign((size_t)n, nullptr);
    int rc = B200RAG_OK;
    for (int i = 0; i < n && rc == B200RAG_OK; ++i) {
        Shard* s = g->shards[(size_t)i];
        cudaError_t e = cudaSetDevice(s->cfg.device);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev[(size_t)i], cudaEventDisableTiming);
        if (e != cudaSuccess) { rc = cuda_fail(e, "group_create: event"); break; }
        if (s->cfg.device != root_dev) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, s->cfg.device, root_dev) == cudaSuccess && can) {
                e = cudaDeviceEnablePeerAccess(root_dev, 0);      // direct NVLink copies; without it the driver stages
                if (e != cudaSuccess) cudaGetLastError();         // (already enabled / not supported: copies still work)
            }
        }
    }
    if (rc != B200RAG_OK) {
        for (auto ev : g->ev) if (ev) cudaEventDestroy(ev);
        delete g;
        return rc;
    }
    // one worker per device, in order of first appearance; the root's device comes first and is run by the caller
    for (int i = 0; i < n; ++i) {
        const int dev = g->shards[(size_t)i]->cfg.device;
        GroupWorker* w = nullptr;
        for (auto* x : g->workers) if (x->device == dev) w = x;
        if (w == nullptr) { w = new GroupWorker(); w->g = g; w->device = dev; g->workers.push_back(w); }
        w->shards.push_back(i);
    }
    for (size_t i = 1; i < g->workers.size(); ++i) g->workers[i]->th = std::thread(worker_main, g->workers[i]);
    cudaSetDevice(root_dev);
    *out = (b200rag_group*)g;
    return B200RAG_OK;
}

void b200rag_group_destroy(b200rag_group* gp) {
    Group* g = (Group*)gp;
    if (g == nullptr) return;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
        ++g->gen;
        for (size_t i = 1; i < g->workers.size(); ++i) g->workers[i]->go.store(g->gen, std::memory_order_release);
    }
    g->cv.notify_all();
    for (size_t i = 1; i < g->workers.size(); ++i) if (g->workers[i]->th.joinable()) g->workers[i]->th.join();
    for (auto* w : g->workers) delete w;
    cudaSetDevice(g->shards[0]->cfg.device);
    cudaStreamSynchronize(g->shards[0]->stream);
    g->window.release();
    g->out.release();
    if (g->h_out) cudaFreeHost(g->h_out);
    for (size_t i = 0; i < g->ev.size(); ++i) {
        if (g->ev[i]) { cudaSetDevice(g->shards[i]->cfg.device); cudaEventDestroy(g->ev[i]); }
    }
    delete g;
}

int32_t b200rag_group_size(const b200rag_group* gp) { return gp ? (int32_t)((const Group*)gp)->shards.size() : 0; }

int b200rag_group_search(b200rag_group* gp, const b200rag_query* q, int64_t* out_ids, double* out_scores,
                         int32_t* out_counts) {
    Group* g = (Group*)gp;
    if (g == nullptr || q == nullptr || out_ids == nullptr || out_scores == nullptr || out_counts == nullptr) {
        set_error("group_search: null argument");
        return B200RAG_ERR_INVALID;
    }
    if (q->batch < 1 || q->top_k < 1 || q->top_k > B200RAG_MAX_TOPK || q->mode < 0 || q->mode > 2) {
        set_error("group_search: bad query");
        return B200RAG_ERR_INVALID;
    }
    const int n = (int)g->shards.size();
    const int B = q->batch, K = q->top_k;
    const int nlegs = q->mode == B200RAG_HYBRID ? 2 : 1;
    const int L = q->mode == B200RAG_HYBRID ? 2 * K : K;
    Shard* root = g->shards[0];
    B2_CUDA(cudaSetDevice(root->cfg.device));
    g->q = q;
    g->block_bytes = ((size_t)nlegs * B * L + 1) * sizeof(b200rag_cand);
    if (g->block_bytes > g->slot_bytes) {
        // (no search is in flight: every call returns only after its results were read back)
        g->slot_bytes = (g->block_bytes + 4095) & ~(size_t)4095;
        B2_TRY(g->window.ensure((size_t)n * g->slot_bytes, 0, root->stream));
    }
    const size_t o_sc = (size_t)B * K * 8, o_cnt = 2 * o_sc;
    const size_t out_bytes = o_cnt + (size_t)(B + 1) * 4;
    B2_TRY(g->out.ensure(out_bytes, 0, root->stream));
    if (out_bytes > g->h_out_cap) {
        if (g->h_out) cudaFreeHost(g->h_out);
        g->h_out = nullptr;
        g->h_out_cap = 0;
        B2_CUDA(cudaMallocHost(&g->h_out, out_bytes * 2));
        g->h_out_cap = out_bytes * 2;
    }
    std::vector<int> saved_slack((size_t)n);
    std::vector<char> saved_ex((size_t)n);
    for (int i = 0; i < n; ++i) { saved_slack[(size_t)i] = g->shards[(size_t)i]->slack; saved_ex[(size_t)i] = g->shards[(size_t)i]->exhaustive; }
    int retries = 0, rc = B200RAG_OK;
    bool unresolved = false;
    int slack = root->slack;
    bool exhaustive = root->exhaustive;
    for (;;) {
        int32_t ambiguous = 0;
        rc = group_pass(g, B, K, L, out_bytes, &ambiguous);
        if (rc != B200RAG_OK || ambiguous == 0) break;
        // some shard's slack guard failed: widen on EVERY shard (the decision is global, like in ShardedSearcher)
        const int cur = slack > 0 ? slack : default_slack_g(L);
        if (exhaustive) { unresolved = true; break; }
        if (L + cur >= 3 * B200RAG_MAX_TOPK || retries >= 6) {
            if (!root->exact_fallback) { unresolved = true; break; }
            exhaustive = true;
        } else {
            slack = cur * 2 + L < 3 * B200RAG_MAX_TOPK - L ? cur * 2 + L : 3 * B200RAG_MAX_TOPK - L;
        }
        for (Shard* s : g->shards) { s->slack = slack; s->exhaustive = exhaustive; }
        ++retries;
    }
    for (int i = 0; i < n; ++i) { g->shards[(size_t)i]->slack = saved_slack[(size_t)i]; g->shards[(size_t)i]->exhaustive = saved_ex[(size_t)i] != 0; }
    g->q = nullptr;
    if (rc != B200RAG_OK) return rc;
    if (unresolved) {
        set_error("group_search: the slack guard never cleared and the exhaustive exact pass is disabled");
        return B200RAG_ERR_INEXACT;
    }
    const uint8_t* h = (const uint8_t*)g->h_out;
    memcpy(out_ids, h, (size_t)B * K * 8);
    memcpy(out_scores, h + o_sc, (size_t)B * K * 8);
    memcpy(out_counts, h + o_cnt, (size_t)B * 4);
    g->last_stats = root->stats;
    g->last_stats.retries = retries;
    g->last_stats.kernel_launches = 0;
    for (Shard* s : g->shards) g->last_stats.kernel_launches += s->stats.kernel_launches;
    return B200RAG_OK;
}

int b200rag_group_get_stats(const b200rag_group* gp, b200rag_stats* out) {
    if (gp == nullptr || out == nullptr) { set_error("group_get_stats: null argument"); return B200RAG_ERR_INVALID; }
    *out = ((const Group*)gp)->last_stats;
    return B200RAG_OK;
}

}  // extern "C"
