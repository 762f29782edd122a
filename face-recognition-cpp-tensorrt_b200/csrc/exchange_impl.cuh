// Cross-GPU exchange of per-shard search results over NVLink peer memory, fused with the search's re-rank on the sending side
// and with the final merge on the receiving side.
//
// The reference has no multi-GPU code; this is the "one exchange step" of the gallery-sharded search. Instead of an NCCL all-gather
// followed by a merge kernel:
//   push   the re-rank kernel of the search (append_rerank_kernel / topk_rerank_kernel; for queries recomputed by the exact scan, its
//          merging block) stores each query's k (score, global row) results straight into EVERY peer's mailbox (remote st.global over
//          NVLink / NVSwitch) the moment that query is finished, then publishes a per-query flag (st.release.sys). No extra launch.
//   merge  one small kernel per rank waits for the peers' flags of each query (ld.acquire.sys), merges all shards' candidates by
//          (score desc, global row asc) and writes the final result.
// Mailboxes are cudaMalloc'ed per rank and mapped into the peers with CUDA IPC (one process per GPU); the 64-byte handles travel once,
// at setup, through any host channel (torch.distributed in bench.py). No host synchronisation and no NCCL call in the step, so the
// whole multi-GPU search step is CUDA-graph capturable.
//
// Mailbox layout on every rank:  entry[slot 4][world][nq_max][k_max] {f32 score, i32 pad, i64 idx}  +  flag[4][world][nq_max] u32.
// Pushes and merges are numbered separately (device-resident counters, so a captured graph can be replayed): push number e uses slot
// e & 3 and flag value e; merge number e waits for flag value e in slot e & 3. The host may run the merge of batch i AFTER the search
// (and push) of batch i + 1 ("lag 1": the flag wait of batch i hides behind the scan of batch i + 1). With lag <= 1 a rank is at most
// 3 pushes ahead of a peer's merge — push e follows the local merge e - 2, which needed the peer's push e - 2, which follows the peer's
// merge e - 4 — so four slots never collide.
// A peer that never arrives is reported, not trapped on: after FR_XCHG_TIMEOUT_MS (default 20 s, %globaltimer) the waiting block records
// the failure in a status word the host can read (fr_exchange_status), substitutes (-inf, -1) for the missing shard and carries on.
// (textually included at the end of gallery.cu: it shares that translation unit's kernels and helpers)
#include <memory>

namespace {

__device__ __forceinline__ void st_flag_system(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_flag_system(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Stand-alone push (results already in device memory): the unfused form, used by fr_exchange_merge_dev, by empty shards and by tests.
// local_s == nullptr pushes (-inf, -1): a shard without rows still takes part in the protocol.
__global__ void __launch_bounds__(64) exchange_push_kernel(XPush xp, int nq, int k, const float* __restrict__ local_s,
                                                           const long long* __restrict__ local_i) {
    __shared__ float s_s[kTopkMax];
    __shared__ long long s_i[kTopkMax];
    const int q = blockIdx.x;
    const unsigned int epoch = xpush_epoch(xp, false);
    if (threadIdx.x < k) {
        s_s[threadIdx.x] = local_s ? local_s[static_cast<size_t>(q) * k + threadIdx.x] : -INFINITY;
        s_i[threadIdx.x] = local_s ? local_i[static_cast<size_t>(q) * k + threadIdx.x] : -1;
    }
    __syncthreads();
    xpush_query(xp, epoch, q, k, s_s, s_i);
    xpush_finish(xp, epoch, gridDim.x);
}

// grid = nq blocks of 64 threads: wait for every shard's entries of this query in MY mailbox, merge, write the result.
__global__ void __launch_bounds__(64) exchange_wait_merge_kernel(XPeers peers, int world, int rank, int nq_max, int k_max, int k,
                                                                 unsigned int* state, unsigned long long timeout_ns,
                                                                 float* __restrict__ out_s, long long* __restrict__ out_i) {
    __shared__ float cs[16 * kTopkMax];
    __shared__ long long ci[16 * kTopkMax];
    __shared__ float sel_s[kTopkMax];
    __shared__ long long sel_i[kTopkMax];
    __shared__ float red_s[32];
    __shared__ long long red_i[32];
    __shared__ int red_p[32];
    __shared__ int arrived[16];
    const int q = blockIdx.x;
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(state + 2) + 1u;
    const int slot = epoch & 3;
    if (threadIdx.x < world) {
        const unsigned int* f = peers.flags[rank] + (static_cast<size_t>(slot) * world + threadIdx.x) * nq_max + q;
        int ok = 1;
        if (ld_flag_system(f) != epoch) {
            const unsigned long long t0 = global_timer_ns();
            unsigned int probes = 0;
            while (ld_flag_system(f) != epoch) {
                if ((++probes & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) {
                    ok = 0;
                    atomicMax(state + 4, 1u + static_cast<unsigned int>(threadIdx.x));  // status: 1 + the rank that never arrived
                    break;
                }
            }
        }
        arrived[threadIdx.x] = ok;
    }
    __syncthreads();
    const int count = world * k;
    for (int t = threadIdx.x; t < count; t += blockDim.x) {
        const int r = t / k, j = t % k;
        XEntry e;
        e.score = -INFINITY;
        e.idx = -1;
        if (arrived[r]) e = peers.entries[rank][((static_cast<size_t>(slot) * world + r) * nq_max + q) * k_max + j];
        cs[t] = e.score;
        ci[t] = e.idx;
    }
    __syncthreads();
    block_select(cs, ci, count, k, sel_s, sel_i, red_s, red_i, red_p);
    if (threadIdx.x < k) {
        out_s[static_cast<size_t>(q) * k + threadIdx.x] = sel_s[threadIdx.x];
        out_i[static_cast<size_t>(q) * k + threadIdx.x] = sel_i[threadIdx.x];
    }
    if (threadIdx.x == 0 && atomicAdd(state + 3, 1u) == gridDim.x - 1) {
        state[3] = 0u;
        __threadfence();
        state[2] = epoch;
    }
}

}  // namespace

struct FrExchange {
    int device = 0, world = 1, rank = 0, nq_max = 0, k_max = 0;
    void* base = nullptr;  // my mailbox: entries then flags
    size_t entry_bytes = 0, flag_bytes = 0;
    XPeers peers{};
    std::vector<void*> opened;
    unsigned int* state_dev = nullptr;  // [0] pushes completed [1] push ticket [2] merges completed [3] merge ticket [4] status
    unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
    bool connected = false;
};

namespace {
XPush make_push(const FrExchange* x) {
    XPush p{};
    p.peers = x->peers;
    p.world = x->world;
    p.rank = x->rank;
    p.nq_max = x->nq_max;
    p.k_max = x->k_max;
    p.state = x->state_dev;
    return p;
}
void check_step(const FrExchange* x, int nq, int k) {
    if (!x) throw ArgError{"null exchange"};
    if (!x->connected) throw StateError{"exchange not connected"};
    if (nq < 1 || nq > x->nq_max || k < 1 || k > x->k_max) throw ArgError{"nq/k exceed the exchange's capacity"};
}
void launch_wait_merge(FrExchange* x, int nq, int k, float* scores_dev, long long* idx_dev, cudaStream_t st) {
    NvtxRange nvtx("fr.exchange.wait_merge");
    exchange_wait_merge_kernel<<<nq, 64, 0, st>>>(x->peers, x->world, x->rank, x->nq_max, x->k_max, k, x->state_dev, x->timeout_ns, scores_dev, idx_dev);
    count_launch();
    FRB_CUDA(cudaGetLastError());
}
}  // namespace

extern "C" {

int fr_exchange_create(int device, int world, int rank, int nq_max, int k_max, FrExchange** out) {
    return guarded([&] {
        if (!out) throw ArgError{"out is null"};
        if (world < 1 || world > 16 || rank < 0 || rank >= world) throw ArgError{"bad world/rank (1..16 ranks)"};
        if (nq_max < 1 || k_max < 1 || k_max > FR_TOPK_MAX) throw ArgError{"bad nq_max/k_max"};
        std::unique_ptr<FrExchange> x(new FrExchange());
        use_device(device);
        x->device = device;
        x->world = world;
        x->rank = rank;
        x->nq_max = nq_max;
        x->k_max = k_max;
        x->entry_bytes = sizeof(XEntry) * 4 * world * nq_max * k_max;
        x->flag_bytes = sizeof(unsigned int) * 4 * world * nq_max;
        if (const char* e = std::getenv("FR_XCHG_TIMEOUT_MS")) x->timeout_ns = std::strtoull(e, nullptr, 10) * 1000000ull;
        FRB_CUDA(cudaMalloc(&x->base, x->entry_bytes + x->flag_bytes));
        FRB_CUDA(cudaMemset(x->base, 0, x->entry_bytes + x->flag_bytes));
        FRB_CUDA(cudaMalloc(&x->state_dev, 8 * sizeof(unsigned int)));
        FRB_CUDA(cudaMemset(x->state_dev, 0, 8 * sizeof(unsigned int)));
        FRB_CUDA(cudaDeviceSynchronize());
        *out = x.release();
    });
}

int fr_exchange_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

/* 64-byte CUDA IPC handle of this rank's mailbox: all-gather it across the ranks (any host channel) and pass the
 * world x 64 bytes, rank-major, to fr_exchange_connect. */
int fr_exchange_local_handle(FrExchange* x, void* out_handle) {
    return guarded([&] {
        if (!x || !out_handle) throw ArgError{"null argument"};
        DeviceGuard dg(x->device);
        cudaIpcMemHandle_t h;
        FRB_CUDA(cudaIpcGetMemHandle(&h, x->base));
        std::memcpy(out_handle, &h, sizeof(h));
    });
}

int fr_exchange_connect(FrExchange* x, const void* all_handles) {
    return guarded([&] {
        if (!x || !all_handles) throw ArgError{"null argument"};
        if (x->connected) throw StateError{"exchange already connected"};
        DeviceGuard dg(x->device);
        for (int r = 0; r < x->world; ++r) {
            void* p = nullptr;
            if (r == x->rank) {
                p = x->base;
            } else {
                cudaIpcMemHandle_t h;
                std::memcpy(&h, static_cast<const char*>(all_handles) + static_cast<size_t>(r) * sizeof(h), sizeof(h));
                FRB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
                x->opened.push_back(p);
            }
            x->peers.entries[r] = static_cast<XEntry*>(p);
            x->peers.flags[r] = reinterpret_cast<unsigned int*>(static_cast<char*>(p) + x->entry_bytes);
        }
        x->connected = true;
    });
}

void fr_exchange_destroy(FrExchange* x) {
    if (!x) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (void* p : x->opened) cudaIpcCloseMemHandle(p);
    cudaFree(x->base);
    cudaFree(x->state_dev);
    if (prev >= 0) cudaSetDevice(prev);
    delete x;
}

/* Same-process variant of fr_exchange_connect (single-process multi-GPU hosts such as the reference's app, and tests): `all` holds
 * the world exchange objects, rank-major. Peer access between their devices is enabled here. */
int fr_exchange_connect_local(FrExchange* x, FrExchange* const* all) {
    return guarded([&] {
        if (!x || !all) throw ArgError{"null argument"};
        if (x->connected) throw StateError{"exchange already connected"};
        DeviceGuard dg(x->device);
        for (int r = 0; r < x->world; ++r) {
            if (!all[r] || all[r]->world != x->world || all[r]->rank != r || all[r]->nq_max != x->nq_max || all[r]->k_max != x->k_max)
                throw ArgError{"exchange objects do not form one group"};
            if (all[r]->device != x->device) {
                int can = 0;
                FRB_CUDA(cudaDeviceCanAccessPeer(&can, x->device, all[r]->device));
                if (!can) throw StateError{"no peer access between the devices of the group"};
                cudaError_t e = cudaDeviceEnablePeerAccess(all[r]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FRB_CUDA(e);
                cudaGetLastError();
            }
            x->peers.entries[r] = static_cast<XEntry*>(all[r]->base);
            x->peers.flags[r] = reinterpret_cast<unsigned int*>(static_cast<char*>(all[r]->base) + all[r]->entry_bytes);
        }
        x->connected = true;
    });
}

/* Search + push, fused: fr_gallery_topk_dev whose re-rank kernel also delivers each query's results to every peer's mailbox. The
 * local results still land in local_scores_dev / local_idx_dev. A shard without rows pushes (-inf, -1) (it must still take part). */
int fr_gallery_topk_push_dev(FrGallery* g, FrExchange* x, const float* q_dev, int nq, int k, float* local_scores_dev, int64_t* local_idx_dev,
                             void* stream) {
    return guarded([&] {
        if (!g) throw ArgError{"null gallery"};
        check_step(x, nq, k);
        if (nq > kChunkQ) throw ArgError{"the exchange step takes one chunk of <= 256 queries"};
        if (!q_dev || !local_scores_dev || !local_idx_dev) throw ArgError{"null argument"};
        if (g->device != x->device) throw ArgError{"gallery shard and exchange live on different devices"};
        DeviceGuard dg(g->device);
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g->stream;
        if (g->n == 0) {
            exchange_push_kernel<<<nq, 64, 0, st>>>(make_push(x), nq, k, nullptr, nullptr);
            count_launch();
            FRB_CUDA(cudaGetLastError());
            return;
        }
        g->first_chunk = true;
        if (g->path == FR_PATH_EXACT || (g->path == FR_PATH_AUTO && g->n < kExactMaxRows) || (g->scan == FR_SCAN_F16 && !g->f16_ok)) {
            // exact fp32 scan (tiny shards, FR_PATH_EXACT): no re-rank kernel to fuse with; deliver with the stand-alone push
            topk_chunk(g, q_dev, nq, k, local_scores_dev, reinterpret_cast<long long*>(local_idx_dev), st);
            exchange_push_kernel<<<nq, 64, 0, st>>>(make_push(x), nq, k, local_scores_dev, reinterpret_cast<const long long*>(local_idx_dev));
            count_launch();
            FRB_CUDA(cudaGetLastError());
            return;
        }
        g->push = make_push(x);
        g->push.enabled = 1;
        try {
            topk_chunk(g, q_dev, nq, k, local_scores_dev, reinterpret_cast<long long*>(local_idx_dev), st);
        } catch (...) {
            g->push.enabled = 0;
            throw;
        }
        g->push.enabled = 0;
    });
}

/* The receiving half: wait for every shard's push of the oldest unmerged batch, merge, write nq x k results (global row ids). */
int fr_exchange_wait_merge_dev(FrExchange* x, int nq, int k, float* scores_dev, int64_t* idx_dev, void* stream) {
    return guarded([&] {
        check_step(x, nq, k);
        if (!scores_dev || !idx_dev) throw ArgError{"null output"};
        DeviceGuard dg(x->device);
        launch_wait_merge(x, nq, k, scores_dev, reinterpret_cast<long long*>(idx_dev), static_cast<cudaStream_t>(stream));
    });
}

/* Unfused form (results of an earlier fr_gallery_topk_dev already in device memory): push + wait + merge, two small launches.
 * local_scores_dev == NULL: this rank has nothing (empty shard) and pushes (-inf, -1). All ranks must issue the same sequence of calls. */
int fr_exchange_merge_dev(FrExchange* x, const float* local_scores_dev, const int64_t* local_idx_dev, int nq, int k, float* scores_dev,
                          int64_t* idx_dev, void* stream) {
    return guarded([&] {
        check_step(x, nq, k);
        if (!scores_dev || !idx_dev || (local_scores_dev && !local_idx_dev)) throw ArgError{"null argument"};
        DeviceGuard dg(x->device);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        exchange_push_kernel<<<nq, 64, 0, st>>>(make_push(x), nq, k, local_scores_dev, reinterpret_cast<const long long*>(local_idx_dev));
        count_launch();
        launch_wait_merge(x, nq, k, scores_dev, reinterpret_cast<long long*>(idx_dev), st);
    });
}

/* 0 = every wait so far was satisfied; r + 1 = a block gave up waiting for rank r (results then lack that shard). Synchronises the device. */
int fr_exchange_status(FrExchange* x, int* out) {
    return guarded([&] {
        if (!x || !out) throw ArgError{"null argument"};
        DeviceGuard dg(x->device);
        FRB_CUDA(cudaDeviceSynchronize());
        unsigned int v = 0;
        FRB_CUDA(cudaMemcpy(&v, x->state_dev + 4, sizeof(v), cudaMemcpyDeviceToHost));
        *out = static_cast<int>(v);
    });
}

/* The public host-buffer search call of a (possibly sharded) gallery — what each rank's host code calls per query batch:
 * queries from host memory -> this shard's fused search (+ push) -> cross-GPU merge -> results to host memory, synchronised.
 * x == NULL: single shard (= fr_gallery_topk with an explicit stream). stream: cudaStream_t as void*, NULL = the gallery's own. */
int fr_search_topk(FrGallery* g, FrExchange* x, const float* q, int nq, int k, float* scores, int64_t* idx, void* stream) {
    return guarded([&] {
        if (!g || !q || !scores || !idx) throw ArgError{"null argument"};
        if (nq < 1 || nq > kChunkQ) throw ArgError{"1..256 queries per call"};
        if (k < 1 || k > FR_TOPK_MAX) throw ArgError{"k out of range"};
        if (!x && g->n == 0) throw StateError{"Feature matching: No faces in database or no faces found"};
        DeviceGuard dg(g->device);
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g->stream;
        FRB_CUDA(cudaMemcpyAsync(g->q_dev, q, sizeof(float) * nq * kDim, cudaMemcpyHostToDevice, st));
        float* fs = g->res_s;
        long long* fi = g->res_i;
        if (x) {
            const int rc = fr_gallery_topk_push_dev(g, x, g->q_dev, nq, k, g->loc_s, reinterpret_cast<int64_t*>(g->loc_i), st);
            if (rc != FR_OK) throw CudaError{std::string("sharded search failed: ") + fr_last_error()};
            launch_wait_merge(x, nq, k, fs, fi, st);
        } else {
            g->first_chunk = true;
            topk_chunk(g, g->q_dev, nq, k, fs, fi, st);
        }
        FRB_CUDA(cudaMemcpyAsync(scores, fs, sizeof(float) * nq * k, cudaMemcpyDeviceToHost, st));
        FRB_CUDA(cudaMemcpyAsync(idx, fi, sizeof(long long) * nq * k, cudaMemcpyDeviceToHost, st));
        FRB_CUDA(cudaStreamSynchronize(st));
    });
}

/* ---- asynchronous host-buffer search stream: the serving form of fr_search_topk. Up to two batches are in flight, so the GPU always
 * has the next search queued while the host collects the previous result, and (sharded) the cross-GPU merge of batch i is issued
 * after the search of batch i + 1 (lag 1). Queries are staged through library-owned pinned memory: the caller's buffers are free again
 * when submit returns. All ranks of a sharded gallery must submit / collect the same batches in the same order. */
}  // extern "C"

struct FrSearchStream {
    static constexpr int kRing = 4;
    FrGallery* g = nullptr;
    FrExchange* x = nullptr;
    int k = 1;
    static constexpr int kMaxInFlight = 3;
    cudaStream_t st = nullptr;       // searches, pushes, merges
    cudaStream_t cp = nullptr;       // H2D of the queries / D2H of the results: overlap the scans of the neighbouring batches
    cudaEvent_t h2d[kRing] = {};     // queries of the slot are on the device
    cudaEvent_t res[kRing] = {};     // results of the slot are final on the device
    float* q_pin[kRing] = {};
    float* q_dev[kRing] = {};
    float* s_pin[kRing] = {};
    long long* i_pin[kRing] = {};
    float* res_s[kRing] = {};
    long long* res_i[kRing] = {};
    cudaEvent_t done[kRing] = {};
    int nq[kRing] = {};
    bool delivered[kRing] = {};     // merge + D2H + event of this batch already enqueued
    long long submitted = 0, collected = 0;
};

namespace {
// enqueue the receiving half of batch b: (sharded) wait + merge, then D2H into pinned memory, then the event collect waits on
void enqueue_delivery(FrSearchStream* s, long long b) {
    const int r = static_cast<int>(b % FrSearchStream::kRing);
    if (s->delivered[r]) return;
    if (s->x) launch_wait_merge(s->x, s->nq[r], s->k, s->res_s[r], s->res_i[r], s->st);
    FRB_CUDA(cudaEventRecord(s->res[r], s->st));
    FRB_CUDA(cudaStreamWaitEvent(s->cp, s->res[r], 0));
    FRB_CUDA(cudaMemcpyAsync(s->s_pin[r], s->res_s[r], sizeof(float) * s->nq[r] * s->k, cudaMemcpyDeviceToHost, s->cp));
    FRB_CUDA(cudaMemcpyAsync(s->i_pin[r], s->res_i[r], sizeof(long long) * s->nq[r] * s->k, cudaMemcpyDeviceToHost, s->cp));
    FRB_CUDA(cudaEventRecord(s->done[r], s->cp));
    s->delivered[r] = true;
}
}  // namespace

extern "C" {

int fr_search_stream_create(FrGallery* g, FrExchange* x, int k, FrSearchStream** out) {
    return guarded([&] {
        if (!g || !out) throw ArgError{"null argument"};
        if (k < 1 || k > FR_TOPK_MAX) throw ArgError{"k out of range"};
        if (x && (!x->connected || x->device != g->device || k > x->k_max)) throw ArgError{"exchange not usable with this gallery / k"};
        std::unique_ptr<FrSearchStream> s(new FrSearchStream());
        s->g = g;
        s->x = x;
        s->k = k;
        DeviceGuard dg(g->device);
        try {
            FRB_CUDA(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
            FRB_CUDA(cudaStreamCreateWithFlags(&s->cp, cudaStreamNonBlocking));
            for (int r = 0; r < FrSearchStream::kRing; ++r) {
                FRB_CUDA(cudaEventCreateWithFlags(&s->h2d[r], cudaEventDisableTiming));
                FRB_CUDA(cudaEventCreateWithFlags(&s->res[r], cudaEventDisableTiming));
                FRB_CUDA(cudaMallocHost(&s->q_pin[r], sizeof(float) * kChunkQ * kDim));
                FRB_CUDA(cudaMalloc(&s->q_dev[r], sizeof(float) * kChunkQ * kDim));
                FRB_CUDA(cudaMallocHost(&s->s_pin[r], sizeof(float) * kChunkQ * FR_TOPK_MAX));
                FRB_CUDA(cudaMallocHost(&s->i_pin[r], sizeof(long long) * kChunkQ * FR_TOPK_MAX));
                FRB_CUDA(cudaMalloc(&s->res_s[r], sizeof(float) * kChunkQ * FR_TOPK_MAX));
                FRB_CUDA(cudaMalloc(&s->res_i[r], sizeof(long long) * kChunkQ * FR_TOPK_MAX));
                FRB_CUDA(cudaEventCreateWithFlags(&s->done[r], cudaEventDisableTiming));
            }
        } catch (...) {
            fr_search_stream_destroy(s.release());
            throw;
        }
        *out = s.release();
    });
}

void fr_search_stream_destroy(FrSearchStream* s) {
    if (!s) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->g->device);
    if (s->st) cudaStreamSynchronize(s->st);
    if (s->cp) cudaStreamSynchronize(s->cp);
    for (int r = 0; r < FrSearchStream::kRing; ++r) {
        if (s->h2d[r]) cudaEventDestroy(s->h2d[r]);
        if (s->res[r]) cudaEventDestroy(s->res[r]);
        cudaFreeHost(s->q_pin[r]);
        cudaFree(s->q_dev[r]);
        cudaFreeHost(s->s_pin[r]);
        cudaFreeHost(s->i_pin[r]);
        cudaFree(s->res_s[r]);
        cudaFree(s->res_i[r]);
        if (s->done[r]) cudaEventDestroy(s->done[r]);
    }
    if (s->st) cudaStreamDestroy(s->st);
    if (s->cp) cudaStreamDestroy(s->cp);
    if (prev >= 0) cudaSetDevice(prev);
    delete s;
}

/* the CUDA stream the searches run on (cudaStream_t as void*): for callers that time the stream with their own events */
void* fr_search_stream_cuda_stream(FrSearchStream* s) { return s ? static_cast<void*>(s->st) : nullptr; }

int fr_search_stream_submit(FrSearchStream* s, const float* q, int nq) {
    return guarded([&] {
        if (!s || !q) throw ArgError{"null argument"};
        if (nq < 1 || nq > kChunkQ) throw ArgError{"1..256 queries per batch"};
        if (s->submitted - s->collected >= FrSearchStream::kMaxInFlight) throw StateError{"three batches already in flight: collect one first"};
        FrGallery* g = s->g;
        if (!s->x && g->n == 0) throw StateError{"Feature matching: No faces in database or no faces found"};
        DeviceGuard dg(g->device);
        const long long b = s->submitted;
        const int r = static_cast<int>(b % FrSearchStream::kRing);
        std::memcpy(s->q_pin[r], q, sizeof(float) * nq * kDim);
        s->nq[r] = nq;
        s->delivered[r] = false;
        // the slot's previous batch (kRing >= kMaxInFlight + 1 batches ago) was collected, so its search has long read q_dev[r]
        FRB_CUDA(cudaMemcpyAsync(s->q_dev[r], s->q_pin[r], sizeof(float) * nq * kDim, cudaMemcpyHostToDevice, s->cp));
        FRB_CUDA(cudaEventRecord(s->h2d[r], s->cp));
        FRB_CUDA(cudaStreamWaitEvent(s->st, s->h2d[r], 0));
        if (s->x) {
            const int rc = fr_gallery_topk_push_dev(g, s->x, s->q_dev[r], nq, s->k, g->loc_s, reinterpret_cast<int64_t*>(g->loc_i), s->st);
            if (rc != FR_OK) throw CudaError{std::string("sharded search failed: ") + fr_last_error()};
        } else {
            g->first_chunk = true;
            topk_chunk(g, s->q_dev[r], nq, s->k, s->res_s[r], s->res_i[r], s->st);
        }
        s->submitted = b + 1;
        // receiving half: single shard -> right away; sharded -> the PREVIOUS batch now (its wait hides behind this batch's scan)
        if (!s->x) enqueue_delivery(s, b);
        else if (b > s->collected) enqueue_delivery(s, b - 1);
    });
}

int fr_search_stream_collect(FrSearchStream* s, float* scores, int64_t* idx, int* nq_out) {
    return guarded([&] {
        if (!s || !scores || !idx) throw ArgError{"null argument"};
        if (s->collected == s->submitted) throw StateError{"nothing in flight"};
        DeviceGuard dg(s->g->device);
        const long long b = s->collected;
        const int r = static_cast<int>(b % FrSearchStream::kRing);
        enqueue_delivery(s, b);  // no newer batch followed: deliver now
        FRB_CUDA(cudaEventSynchronize(s->done[r]));
        std::memcpy(scores, s->s_pin[r], sizeof(float) * s->nq[r] * s->k);
        std::memcpy(idx, s->i_pin[r], sizeof(long long) * s->nq[r] * s->k);
        if (nq_out) *nq_out = s->nq[r];
        s->collected = b + 1;
    });
}

}  // extern "C"
