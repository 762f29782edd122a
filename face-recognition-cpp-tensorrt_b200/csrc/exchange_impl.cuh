// Cross-GPU exchange of per-shard search results over NVLink peer memory, fused with the final merge.
//
// The reference has no multi-GPU code; this is the "one exchange step" of the gallery-sharded search. Instead of an NCCL
// all-gather followed by a merge kernel, ONE kernel per rank (a) stores its nq x k (score, idx) results straight into every
// peer's mailbox (remote st.global over NVLink / NVSwitch), (b) publishes a per-query flag, (c) waits for the peers' flags and
// (d) merges all shards' candidates by (score desc, global row asc). Mailboxes are cudaMalloc'ed per rank and mapped into the
// peers with CUDA IPC (one process per GPU); the 64-byte handles travel once, at setup, through any host channel
// (torch.distributed in bench.py). No host synchronisation and no NCCL call in the step, so the whole search step is
// CUDA-graph capturable.
//
// Mailbox layout on every rank:  entry[parity 2][world][nq_max][k_max] {f32 score, i32 pad, i64 idx}  +  flag[2][world][nq_max] u32.
// A call with sequence number `epoch` uses parity = epoch & 1; a rank cannot be two calls ahead of a peer (its merge of call e+1
// needs the peer's push of e+1, which is stream-ordered after the peer's merge of call e), so two parities suffice.
// (textually included at the end of gallery.cu: it shares that translation unit's kernels and helpers)
#include <memory>

namespace {

struct __align__(16) XEntry {
    float score;
    int pad;
    long long idx;
};

struct XPeers {
    XEntry* entries[16];
    unsigned int* flags[16];
};

__device__ __forceinline__ void st_flag_system(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_flag_system(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The call sequence number lives in device memory so that a CUDA graph of the step can be replayed. Every block of a call reads
// epoch_state[0] + 1 when it starts; the block that finishes last (ticket counter epoch_state[1]) stores the new value, i.e. after
// every block of the call has read the old one and before the next call (stream order) starts.

// grid = nq blocks of 64 threads. local_s / local_i: this rank's nq x k results (global row ids).
__global__ void __launch_bounds__(64) exchange_merge_kernel(XPeers peers, int world, int rank, int nq_max, int k_max, int nq, int k,
                                                            unsigned int* epoch_state, const float* __restrict__ local_s,
                                                            const long long* __restrict__ local_i, float* __restrict__ out_s,
                                                            long long* __restrict__ out_i) {
    __shared__ float cs[16 * kTopkMax];
    __shared__ long long ci[16 * kTopkMax];
    __shared__ float sel_s[kTopkMax];
    __shared__ long long sel_i[kTopkMax];
    __shared__ float red_s[32];
    __shared__ long long red_i[32];
    __shared__ int red_p[32];
    const int q = blockIdx.x;
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(epoch_state) + 1u;
    const int par = epoch & 1;
    const size_t slot = ((static_cast<size_t>(par) * world + rank) * nq_max + q) * k_max;  // my slot in every mailbox
    // (a) push: thread t -> (peer t / k, entry t % k)
    for (int t = threadIdx.x; t < world * k; t += blockDim.x) {
        const int peer = t / k, j = t % k;
        XEntry e;
        e.score = local_s[static_cast<size_t>(q) * k + j];
        e.pad = 0;
        e.idx = local_i[static_cast<size_t>(q) * k + j];
        peers.entries[peer][slot + j] = e;
    }
    __threadfence_system();
    __syncthreads();
    // (b) publish
    if (threadIdx.x < world)
        st_flag_system(peers.flags[threadIdx.x] + (static_cast<size_t>(par) * world + rank) * nq_max + q, epoch);
    // (c) wait for every shard's entry of this query in MY mailbox
    if (threadIdx.x < world) {
        const unsigned int* f = peers.flags[rank] + (static_cast<size_t>(par) * world + threadIdx.x) * nq_max + q;
        unsigned int spins = 0;
        while (ld_flag_system(f) != epoch) {
            if (++spins == (1u << 25)) __trap();  // a peer never arrived: fail loudly instead of hanging the GPU
        }
    }
    __syncthreads();
    // (d) merge
    const int count = world * k;
    for (int t = threadIdx.x; t < count; t += blockDim.x) {
        const int r = t / k, j = t % k;
        const XEntry e = peers.entries[rank][((static_cast<size_t>(par) * world + r) * nq_max + q) * k_max + j];
        cs[t] = e.score;
        ci[t] = e.idx;
    }
    __syncthreads();
    block_select(cs, ci, count, k, sel_s, sel_i, red_s, red_i, red_p);
    if (threadIdx.x < k) {
        out_s[static_cast<size_t>(q) * k + threadIdx.x] = sel_s[threadIdx.x];
        out_i[static_cast<size_t>(q) * k + threadIdx.x] = sel_i[threadIdx.x];
    }
    if (threadIdx.x == 0 && atomicAdd(epoch_state + 1, 1u) == gridDim.x - 1) {
        epoch_state[1] = 0u;
        epoch_state[0] = epoch;
    }
}

}  // namespace

struct FrExchange {
    int device = 0, world = 1, rank = 0, nq_max = 0, k_max = 0;
    void* base = nullptr;  // my mailbox: entries then flags
    size_t entry_bytes = 0, flag_bytes = 0;
    XPeers peers{};
    std::vector<void*> opened;
    unsigned int* epoch_dev = nullptr;
    bool connected = false;
};

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
        x->entry_bytes = sizeof(XEntry) * 2 * world * nq_max * k_max;
        x->flag_bytes = sizeof(unsigned int) * 2 * world * nq_max;
        FRB_CUDA(cudaMalloc(&x->base, x->entry_bytes + x->flag_bytes));
        FRB_CUDA(cudaMemset(x->base, 0, x->entry_bytes + x->flag_bytes));
        FRB_CUDA(cudaMalloc(&x->epoch_dev, 2 * sizeof(unsigned int)));  // [0] calls completed, [1] finished-block ticket of the running call
        FRB_CUDA(cudaMemset(x->epoch_dev, 0, 2 * sizeof(unsigned int)));
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
    cudaFree(x->epoch_dev);
    if (prev >= 0) cudaSetDevice(prev);
    delete x;
}

/* The exchange step: every rank calls it once per search with its own nq x k results (device, global row ids, as written by
 * fr_gallery_topk_dev); on return (stream order) scores_dev / idx_dev hold the merged global top-k on every rank.
 * All ranks must issue the same sequence of calls. Does not synchronise the host. */
int fr_exchange_merge_dev(FrExchange* x, const float* local_scores_dev, const int64_t* local_idx_dev, int nq, int k, float* scores_dev,
                          int64_t* idx_dev, void* stream) {
    return guarded([&] {
        if (!x || !local_scores_dev || !local_idx_dev || !scores_dev || !idx_dev) throw ArgError{"null argument"};
        if (!x->connected) throw StateError{"exchange not connected"};
        if (nq < 1 || nq > x->nq_max || k < 1 || k > x->k_max) throw ArgError{"nq/k exceed the exchange's capacity"};
        DeviceGuard dg(x->device);
        exchange_merge_kernel<<<nq, 64, 0, static_cast<cudaStream_t>(stream)>>>(x->peers, x->world, x->rank, x->nq_max, x->k_max, nq, k, x->epoch_dev,
                                                                               local_scores_dev,
                                                                               reinterpret_cast<const long long*>(local_idx_dev), scores_dev,
                                                                               reinterpret_cast<long long*>(idx_dev));
        count_launch();
        FRB_CUDA(cudaGetLastError());
    });
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

}  // extern "C"
