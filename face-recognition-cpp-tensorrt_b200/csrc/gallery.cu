// Gallery / cosine-similarity search: host side of the C ABI (include/fr_b200.h, "Gallery" section).
// Replaces MatMul (/root/reference src/matmul.{h,cpp}) and the host argmax ArcFaceIR50::getOutputs (src/arcface.cpp:203-217).
//
// Resident layout per gallery (one shard of a row-partitioned gallery):
//   rows_f32  n x 512 f32 row-major   master copy: dense sims, exact re-score and exact scan read it     (2 KiB / row)
//   rows_f16  n x 512 f16 row-major   scan copy streamed by the fused tensor-core kernel through TMA     (1 KiB / row)
//   gmax      largest row norm (device scalar): scales the provable error margin of the fp16 scan
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "common.h"
#include "search_kernels.cuh"

using namespace frb;

struct FrGallery {
    int device = 0;
    int sms = 0;
    int64_t n = 0;
    int64_t capacity = 0;            // rows the device buffers can hold (>= n); grown by fr_gallery_reserve / fr_gallery_append
    int64_t row_offset = 0;
    float* rows_f32 = nullptr;
    __half* rows_f16 = nullptr;
    uint8_t* rows_f8 = nullptr;       // optional e4m3 scan copy (FR_SCAN_F8), 512 B / row
    float* gmax = nullptr;
    float* amax = nullptr;            // largest |component| of any row
    bool f16_ok = true;               // rows fit fp16's normal range (amax <= 65504, gmax >= kF16MinNorm): else searches take the exact scan
    float* g4max = nullptr;           // largest sum of fourth powers of a row (set with the e4m3 copy): scales the fp8 margin
    float* w4max = nullptr;           // largest sum of fourth powers of a row's e4m3 rounding steps (same)
    uint64_t f8_seed = 0x5EEDF8B200ull;  // dither stream of the stochastic e4m3 rounding (FR_F8_SEED overrides)
    CUtensorMap tmap{};
    CUtensorMap tmap8{};
    int scan = FR_SCAN_F16;
    cudaStream_t stream = nullptr;
    // scratch, sized for one chunk of 256 queries
    float* q_dev = nullptr;          // 256 x 512
    void* q_img = nullptr;           // 256 x 512 fp16 (or e4m3): the scan's query operand, written by prep_queries_kernel per search
    float* q_margin = nullptr;       // 256: scan margin in accumulator units (fp16: 2 eps |q| gmax; e4m3: (1 + gap) E, search_kernels.cuh)
    float* q_gap = nullptr;          // 256: certificate gap of the e4m3 scan in cosine units (+inf on the fp16 copy)
    bool first_chunk = true;         // the first chunk of a topk call resets flagged_acc (inside its prep kernel: no extra launch)
    int* flagged_acc = nullptr;      // device: flagged queries accumulated over the chunks of the last topk call
    CUtensorMap tmq16{}, tmq8{};     // two views of q_img
    float* cand_s = nullptr;         // [lists <= 296][256][16]
    int* cand_i = nullptr;
    int* gbest = nullptr;            // 256: best coarse score per query shared by the scan's epilogue threads (0 between searches)
    uint2* app_buf = nullptr;        // append epilogue: [lists <= 296][256][kAppCap] (coarse score bits, local row), allocated on first use
    int* app_cnt = nullptr;          // [lists][256] entries appended
    int* flags = nullptr;            // [0] = count, [1..256] = queries handed to the exact scan
    unsigned int* scan_ticket = nullptr;  // finished-block counter of the exact-scan fix-up (0 between launches)
    float* part_s = nullptr;         // exact scan partials [256][slices][8]
    long long* part_i = nullptr;
    float* res_s = nullptr;          // 256 x FR_TOPK_MAX
    long long* res_i = nullptr;
    float* loc_s = nullptr;          // this shard's results inside fr_search_topk (sharded), 256 x FR_TOPK_MAX
    long long* loc_i = nullptr;
    XPush push{};                    // enabled only inside fr_gallery_topk_push_dev: the re-rank kernels then also deliver to the peers
    float* sims_ws = nullptr;        // dense-path workspace
    size_t sims_ws_floats = 0;
    int path = FR_PATH_AUTO;
    FrSearchStats stats{};
    // optional CUDA-event timing of the dominant (fused scan) kernel, for bench.py's roofline
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    size_t ev_used = 0;
    // timing mode 2: ONE fixed event pair per scan copy around the fused kernel. Fixed events survive stream capture: a replayed CUDA
    // graph re-records them, so bench.py reads the kernel's duration from the very replays it times (fr_gallery_last_scan_ms).
    bool timing_fixed = false;
    cudaEvent_t fixed_ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
};

namespace {

constexpr int kChunkQ = 256;
constexpr int kMaxLists = 2 * 148;
constexpr int64_t kExactMaxRows = 2048;  // FR_PATH_AUTO: below this the exact scan has the lower latency

void alloc_common(FrGallery* g) {
    FRB_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    FRB_CUDA(cudaMalloc(&g->q_dev, sizeof(float) * kChunkQ * kDim));
    FRB_CUDA(cudaMalloc(&g->q_img, sizeof(__half) * kChunkQ * kDim));
    FRB_CUDA(cudaMalloc(&g->q_margin, sizeof(float) * kChunkQ));
    FRB_CUDA(cudaMalloc(&g->q_gap, sizeof(float) * kChunkQ));
    FRB_CUDA(cudaMalloc(&g->flagged_acc, sizeof(int)));
    FRB_CUDA(cudaMemsetAsync(g->flagged_acc, 0, sizeof(int), g->stream));
    if (const char* e = std::getenv("FR_F8_SEED")) g->f8_seed = std::strtoull(e, nullptr, 0);
    g->tmq16 = make_tmap_2d_f16(g->q_img, kChunkQ, kDim, 128, 64);
    g->tmq8 = make_tmap_2d_u8(g->q_img, kChunkQ, kDim, 128, 128);
    FRB_CUDA(cudaMalloc(&g->cand_s, sizeof(float) * kMaxLists * kChunkQ * 16));
    FRB_CUDA(cudaMalloc(&g->cand_i, sizeof(int) * kMaxLists * kChunkQ * 16));
    FRB_CUDA(cudaMalloc(&g->flags, sizeof(int) * (kChunkQ + 1)));
    FRB_CUDA(cudaMalloc(&g->scan_ticket, sizeof(unsigned int)));
    FRB_CUDA(cudaMemsetAsync(g->scan_ticket, 0, sizeof(unsigned int), g->stream));
    FRB_CUDA(cudaMalloc(&g->gbest, sizeof(int) * kChunkQ));
    FRB_CUDA(cudaMemsetAsync(g->gbest, 0, sizeof(int) * kChunkQ, g->stream));
    FRB_CUDA(cudaMalloc(&g->part_s, sizeof(float) * kChunkQ * kScanSlicesMax * kTopkMax));
    FRB_CUDA(cudaMalloc(&g->part_i, sizeof(long long) * kChunkQ * kScanSlicesMax * kTopkMax));
    FRB_CUDA(cudaMalloc(&g->res_s, sizeof(float) * kChunkQ * FR_TOPK_MAX));
    FRB_CUDA(cudaMalloc(&g->res_i, sizeof(long long) * kChunkQ * FR_TOPK_MAX));
    FRB_CUDA(cudaMalloc(&g->loc_s, sizeof(float) * kChunkQ * FR_TOPK_MAX));
    FRB_CUDA(cudaMalloc(&g->loc_i, sizeof(long long) * kChunkQ * FR_TOPK_MAX));
    FRB_CUDA(cudaMalloc(&g->gmax, sizeof(float)));
    FRB_CUDA(cudaMemsetAsync(g->gmax, 0, sizeof(float), g->stream));
    FRB_CUDA(cudaMalloc(&g->g4max, sizeof(float)));
    FRB_CUDA(cudaMemsetAsync(g->g4max, 0, sizeof(float), g->stream));
    FRB_CUDA(cudaMalloc(&g->amax, sizeof(float)));
    FRB_CUDA(cudaMemsetAsync(g->amax, 0, sizeof(float), g->stream));
    FRB_CUDA(cudaMalloc(&g->w4max, sizeof(float)));
    FRB_CUDA(cudaMemsetAsync(g->w4max, 0, sizeof(float), g->stream));
}

FrGallery* new_gallery(int64_t n, int dim, int device, int64_t row_offset) {
    if (dim != kDim) throw ArgError{"dim must be 512 (rec_outputDim)"};
    if (n < 0 || n >= (int64_t(1) << 31) - kTileRows) throw ArgError{"row count out of range for one shard"};
    auto* g = new FrGallery();
    g->sms = use_device(device);
    g->device = device;
    g->n = n;
    g->capacity = n;
    g->row_offset = row_offset;
    try {
        alloc_common(g);
        if (n > 0) {
            FRB_CUDA(cudaMalloc(&g->rows_f32, sizeof(float) * n * kDim));
            FRB_CUDA(cudaMalloc(&g->rows_f16, sizeof(__half) * n * kDim));
            g->tmap = make_tmap_2d_f16(g->rows_f16, static_cast<uint64_t>(n), kDim, 128, 64);
        }
    } catch (...) {
        fr_gallery_destroy(g);
        throw;
    }
    return g;
}

// the fp16 scan copy's error bound needs rows in fp16's normal range (search_kernels.cuh, kF16MinNorm); the caller has synchronised
void refresh_f16_ok(FrGallery* g) {
    float b[2] = {0.f, 0.f};
    FRB_CUDA(cudaMemcpy(&b[0], g->gmax, sizeof(float), cudaMemcpyDeviceToHost));
    FRB_CUDA(cudaMemcpy(&b[1], g->amax, sizeof(float), cudaMemcpyDeviceToHost));
    g->f16_ok = g->n == 0 || (b[1] <= 65504.f && b[0] >= kF16MinNorm);
}

// fp16 scan copy (unless the generator already wrote it) + largest row norm
void finish_rows(FrGallery* g, bool write_f16) {
    if (g->n == 0) return;
    const int blocks = static_cast<int>(std::min<int64_t>((g->n + 7) / 8, g->sms * 16LL));
    make_scan_copy_kernel<<<blocks, 256, 0, g->stream>>>(g->rows_f32, write_f16 ? g->rows_f16 : nullptr, g->n, g->gmax, g->amax);
    count_launch();
    FRB_CUDA(cudaGetLastError());
    FRB_CUDA(cudaStreamSynchronize(g->stream));
    refresh_f16_ok(g);
}

// tensor maps follow the row count (rows past n are TMA zero fill) and the buffers' base addresses
void refresh_tmaps(FrGallery* g) {
    if (g->n > 0 && g->rows_f16) g->tmap = make_tmap_2d_f16(g->rows_f16, static_cast<uint64_t>(g->n), kDim, 128, 64);
    if (g->n > 0 && g->rows_f8) g->tmap8 = make_tmap_2d_u8(g->rows_f8, static_cast<uint64_t>(g->n), kDim, 128, 128);
}

// grow the resident copies to `capacity` rows, preserving the first n
void grow_rows(FrGallery* g, int64_t capacity) {
    if (capacity <= g->capacity) return;
    if (capacity >= (int64_t(1) << 31) - kTileRows) throw ArgError{"row count out of range for one shard"};
    float* f32 = nullptr;
    __half* f16 = nullptr;
    uint8_t* f8 = nullptr;
    try {
        FRB_CUDA(cudaMalloc(&f32, sizeof(float) * capacity * kDim));
        FRB_CUDA(cudaMalloc(&f16, sizeof(__half) * capacity * kDim));
        if (g->rows_f8 || g->scan == FR_SCAN_F8) FRB_CUDA(cudaMalloc(&f8, static_cast<size_t>(capacity) * kDim));
        if (g->n > 0) {
            FRB_CUDA(cudaMemcpyAsync(f32, g->rows_f32, sizeof(float) * g->n * kDim, cudaMemcpyDeviceToDevice, g->stream));
            FRB_CUDA(cudaMemcpyAsync(f16, g->rows_f16, sizeof(__half) * g->n * kDim, cudaMemcpyDeviceToDevice, g->stream));
            if (f8 && g->rows_f8) FRB_CUDA(cudaMemcpyAsync(f8, g->rows_f8, static_cast<size_t>(g->n) * kDim, cudaMemcpyDeviceToDevice, g->stream));
        }
        FRB_CUDA(cudaStreamSynchronize(g->stream));
    } catch (...) {
        cudaFree(f32);
        cudaFree(f16);
        cudaFree(f8);
        throw;
    }
    cudaFree(g->rows_f32);
    cudaFree(g->rows_f16);
    cudaFree(g->rows_f8);
    g->rows_f32 = f32;
    g->rows_f16 = f16;
    g->rows_f8 = f8;
    g->capacity = capacity;
    refresh_tmaps(g);
}

// FR_F8_LOGP overrides ln(1/p) of the e4m3 certificate (p = per-query bound on a wrong top-1; experiments); FR_SEARCH_APPEND: 0 =
// sorted register lists everywhere, 1 = append epilogue for top-1 searches on the fp8 scan copy only, 2 (default) = on either scan copy
float f8_logp() {
    static const float v = std::getenv("FR_F8_LOGP") ? static_cast<float>(std::atof(std::getenv("FR_F8_LOGP"))) : kF8LogP;
    return v;
}
int append_mode() {
    static const int v = std::getenv("FR_SEARCH_APPEND") ? std::atoi(std::getenv("FR_SEARCH_APPEND")) : 2;
    return v;
}

template <int CG, int KSEL, bool F8, bool APP = false>
void launch_coarse(FrGallery* g, const float* q_dev, int nq, int units, int tiles, cudaStream_t st) {
    prep_queries_kernel<F8><<<kChunkQ / 8, 256, 0, st>>>(q_dev, nq, g->gmax, g->g4max, g->w4max, F8 ? f8_logp() : kCoarseEps, g->f8_seed, g->q_img,
                                                         g->q_margin, g->q_gap, g->first_chunk ? g->flagged_acc : nullptr);
    count_launch();
    std::pair<cudaEvent_t, cudaEvent_t>* ev = nullptr;
    std::pair<cudaEvent_t, cudaEvent_t> fixed;
    // inside a stream capture a plain cudaEventRecord only orders captured work; cudaEventRecordExternal makes it an event-record NODE,
    // so every replay of the graph records (and times) the events for real
    unsigned int ev_flags = cudaEventRecordDefault;
    if (g->timing || g->timing_fixed) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        FRB_CUDA(cudaStreamIsCapturing(st, &cs));
        if (cs == cudaStreamCaptureStatusActive) ev_flags = cudaEventRecordExternal;
    }
    if (g->timing_fixed) {
        for (int j = 0; j < 2; ++j)
            if (!g->fixed_ev[F8][j]) FRB_CUDA(cudaEventCreate(&g->fixed_ev[F8][j]));
        fixed = {g->fixed_ev[F8][0], g->fixed_ev[F8][1]};
        ev = &fixed;
        FRB_CUDA(cudaEventRecordWithFlags(ev->first, st, ev_flags));
    } else if (g->timing) {
        if (g->ev_used == g->ev_pool.size()) {
            cudaEvent_t a, b;
            FRB_CUDA(cudaEventCreate(&a));
            FRB_CUDA(cudaEventCreate(&b));
            g->ev_pool.emplace_back(a, b);
        }
        ev = &g->ev_pool[g->ev_used++];
        FRB_CUDA(cudaEventRecordWithFlags(ev->first, st, ev_flags));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(units * CG);
    cfg.blockDim = dim3(coarse_threads<CG, F8, APP>());
    cfg.dynamicSmemBytes = CoarseCfg<CG, F8>::kSmemBytes;
    cfg.stream = st;
    static bool attr_done[16][2][2][2][2] = {};
    bool& done = attr_done[g->device & 15][CG - 1][KSEL == 1][F8][APP];
    if (!done) {
        FRB_CUDA(cudaFuncSetAttribute(cosine_topk_coarse<CG, KSEL, F8, APP>, cudaFuncAttributeMaxDynamicSharedMemorySize, CoarseCfg<CG, F8>::kSmemBytes));
        done = true;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FRB_CUDA(cudaLaunchKernelEx(&cfg, cosine_topk_coarse<CG, KSEL, F8, APP>, F8 ? g->tmap8 : g->tmap, F8 ? g->tmq8 : g->tmq16,
                                static_cast<const float*>(g->q_margin), nq, static_cast<long long>(g->n), tiles, g->cand_s, g->cand_i, g->flags,
                                g->gbest, g->app_buf, g->app_cnt));
    count_launch();
    if (ev) FRB_CUDA(cudaEventRecordWithFlags(ev->second, st, ev_flags));
}

void ensure_sims_ws(FrGallery* g, size_t floats) {
    if (g->sims_ws_floats >= floats) return;
    if (g->sims_ws) cudaFree(g->sims_ws);
    g->sims_ws = nullptr;
    g->sims_ws_floats = 0;
    FRB_CUDA(cudaMalloc(&g->sims_ws, floats * sizeof(float)));
    g->sims_ws_floats = floats;
}

void launch_sims(FrGallery* g, const float* q_dev, int nq, float* out_dev, cudaStream_t st) {
    const int warps_per_block = kSimsThreads / 32;
    const int gx = static_cast<int>(std::min<int64_t>((g->n + warps_per_block - 1) / warps_per_block, g->sms * 8LL));
    dim3 grid(gx, (nq + kSimsQ - 1) / kSimsQ);
    sims_kernel<<<grid, kSimsThreads, 0, st>>>(g->rows_f32, g->n, q_dev, nq, out_dev);
    count_launch();
    FRB_CUDA(cudaGetLastError());
}

// exact scan of the queries flagged in `flags` (nullptr = all): two launches, both return at once for unflagged queries
void launch_exact(FrGallery* g, const float* q_dev, int nq, int k, const int* flags, float* scores_dev, long long* idx_dev,
                  cudaStream_t st) {
    const int slices = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((g->n + 7) / 8, std::min(kScanSlicesMax, 2 * g->sms))));
    // flagged-query fix-up: one pass of blocks that loop over the (normally empty) list; exact path: spread the queries too
    const int qsplit = flags ? 1 : std::min(nq, 64);
    exact_scan_kernel<<<dim3(slices, qsplit), kScanThreads, 0, st>>>(g->rows_f32, g->n, q_dev, nq, flags, g->part_s, g->part_i, k, g->row_offset,
                                                                     scores_dev, idx_dev, g->scan_ticket, flags ? g->flagged_acc : nullptr,
                                                                     flags ? g->push : XPush{});
    count_launch();
    if (!flags) {  // all queries: the merge is spread over its own grid; the fix-up's last block merges in the same launch
        exact_merge_kernel<<<std::min(nq, 148), kSelThreads, 0, st>>>(g->part_s, g->part_i, slices, nq, flags, k, g->row_offset, scores_dev, idx_dev);
        count_launch();
    }
    FRB_CUDA(cudaGetLastError());
}

// one chunk (nq <= 256) of queries already on the device; results to scores_dev / idx_dev (device, nq x k)
void topk_chunk(FrGallery* g, const float* q_dev, int nq, int k, float* scores_dev, long long* idx_dev, cudaStream_t st) {
    NvtxRange nvtx("fr.search.topk");
    // rows outside fp16's normal range void the fp16 copy's error bound (and the e4m3 copy only ever holds unit rows): exact scan
    const bool exact = g->path == FR_PATH_EXACT || (g->path == FR_PATH_AUTO && g->n < kExactMaxRows) || (g->scan == FR_SCAN_F16 && !g->f16_ok);
    g->stats = FrSearchStats{};
    if (exact) {
        if (g->first_chunk) FRB_CUDA(cudaMemsetAsync(g->flagged_acc, 0, sizeof(int), st));  // nothing is ever flagged on this path
        launch_exact(g, q_dev, nq, k, nullptr, scores_dev, idx_dev, st);
        g->stats.scan_bytes = g->n * kDim * 4 * nq;
        g->stats.flops = 2LL * nq * g->n * kDim;
        g->stats.launches = 2;
        g->stats.ctas = 0;
        return;
    }
    const int tiles = static_cast<int>((g->n + kTileRows - 1) / kTileRows);
    const int cg = nq > kQRows ? 2 : 1;
    const int kc = k == 1 ? 8 : 16;
    int units;
    // k > 1 on an e4m3 gallery is served by the resident fp16 copy: the sorted-list epilogue a top-k needs overflows under the e4m3
    // copy's wide certified margin (unmatched queries fell through to the exact scan: correct but slow), while the fp16 copy's margin is
    // a few 1e-3 and its result deterministic. The e4m3 copy keeps the top-1 searches (getOutputs, src/arcface.cpp:203-217).
    // (FR_F8_TOPK=1, read per call on this rare path, keeps k > 1 on the e4m3 copy: the accumulation test reads its coarse scores)
    const bool f8 = g->scan == FR_SCAN_F8 && (k == 1 || !g->f16_ok || [] { const char* e = std::getenv("FR_F8_TOPK"); return e && e[0] == '1'; }());
    const bool app = k == 1 && (append_mode() >= 2 || (append_mode() == 1 && f8));
    if (app && !g->app_buf) {
        // not reached while a stream capture is open: callers warm a search up before capturing it (cudaMalloc is not capturable)
        FRB_CUDA(cudaMalloc(&g->app_buf, sizeof(uint2) * kMaxLists * kChunkQ * kAppCap));
        FRB_CUDA(cudaMalloc(&g->app_cnt, sizeof(int) * kMaxLists * kChunkQ));
    }
    if (cg == 2) {
        // scratch and the re-rank kernels are sized for kMaxLists lists (the e4m3 append scan writes four lists per unit)
        units = std::min({g->sms / 2, tiles, kMaxLists / (f8 && app ? coarse_epi_warps<2, true, true>() / 4 : 2)});
        if (f8) {
            if (app) launch_coarse<2, 1, true, true>(g, q_dev, nq, units, tiles, st);
            else if (k == 1) launch_coarse<2, 1, true>(g, q_dev, nq, units, tiles, st);
            else launch_coarse<2, 8, true>(g, q_dev, nq, units, tiles, st);
        } else {
            if (app) launch_coarse<2, 1, false, true>(g, q_dev, nq, units, tiles, st);
            else if (k == 1) launch_coarse<2, 1, false>(g, q_dev, nq, units, tiles, st);
            else launch_coarse<2, 8, false>(g, q_dev, nq, units, tiles, st);
        }
    } else {
        units = std::min({g->sms, tiles, kMaxLists / 2});
        if (f8) {
            if (app) launch_coarse<1, 1, true, true>(g, q_dev, nq, units, tiles, st);
            else if (k == 1) launch_coarse<1, 1, true>(g, q_dev, nq, units, tiles, st);
            else launch_coarse<1, 8, true>(g, q_dev, nq, units, tiles, st);
        } else {
            if (app) launch_coarse<1, 1, false, true>(g, q_dev, nq, units, tiles, st);
            else if (k == 1) launch_coarse<1, 1, false>(g, q_dev, nq, units, tiles, st);
            else launch_coarse<1, 8, false>(g, q_dev, nq, units, tiles, st);
        }
    }
    if (app)
        append_rerank_kernel<<<nq, kSelThreads, 0, st>>>(g->app_buf, g->app_cnt, units * (cg == 2 && f8 ? coarse_epi_warps<2, true, true>() / 4 : 2), cg * kQRows, q_dev, g->rows_f32, g->q_margin,
                                                         g->q_gap, f8 ? 1.f / (kF8Scale * kF8Scale) : 1.f, g->row_offset, scores_dev, idx_dev, g->flags,
                                                         g->gbest, g->push, f8 && g->f16_ok ? g->rows_f16 : nullptr, g->gmax);
    else
        topk_rerank_kernel<<<nq, kSelThreads, 0, st>>>(g->cand_s, g->cand_i, units * 2, cg * kQRows, kc, q_dev, g->rows_f32, g->q_margin,
                                                       g->q_gap, f8 ? 1.f / (kF8Scale * kF8Scale) : 1.f, k, g->row_offset, scores_dev, idx_dev, g->flags,
                                                       g->gbest, g->push);
    count_launch();
    FRB_CUDA(cudaGetLastError());
    // queries whose candidate set may be incomplete (flagged by the re-rank) are recomputed exactly; no-op otherwise
    launch_exact(g, q_dev, nq, k, g->flags, scores_dev, idx_dev, st);
    g->stats.scan_bytes = g->n * kDim * (f8 ? 1 : 2);
    g->stats.flops = 2LL * (cg * kQRows) * g->n * kDim;
    g->stats.launches = 4;
    g->stats.ctas = units * cg;
}

// the e4m3 copy scales by 256 and saturates at 448: only L2-normalised rows may enter it
void check_unit_rows(const float* rows, int64_t n) {
    for (int64_t r = 0; r < n; ++r) {
        double ss = 0;
        for (int i = 0; i < kDim; ++i) ss += static_cast<double>(rows[r * kDim + i]) * rows[r * kDim + i];
        if (ss > 1.001 * 1.001) throw StateError{"FR_SCAN_F8 needs L2-normalised rows (row norm > 1)"};
    }
}

void check_query_args(const FrGallery* g, const void* q, int nq) {
    if (!g) throw ArgError{"null gallery"};
    if (!q || nq <= 0) throw ArgError{"no queries"};
    if (g->n == 0) throw StateError{"Feature matching: No faces in database or no faces found"};  // src/arcface.cpp:198
}

void add_stats(FrSearchStats& total, const FrSearchStats& s) {
    total.scan_bytes += s.scan_bytes;
    total.flops += s.flops;
    total.launches += s.launches;
    total.ctas = s.ctas;
}

}  // namespace

extern "C" {

int fr_gallery_create(const float* rows, int64_t n, int dim, int device, int64_t row_offset, FrGallery** out) {
    return guarded([&] {
        if (!out) throw ArgError{"out is null"};
        if (n > 0 && !rows) throw ArgError{"rows is null"};
        FrGallery* g = new_gallery(n, dim, device, row_offset);
        try {
            if (n > 0) {
                FRB_CUDA(cudaMemcpyAsync(g->rows_f32, rows, sizeof(float) * n * kDim, cudaMemcpyHostToDevice, g->stream));
                finish_rows(g, true);
            }
        } catch (...) {
            fr_gallery_destroy(g);
            throw;
        }
        *out = g;
    });
}

int fr_gallery_create_dev(const float* rows_dev, int64_t n, int dim, int device, int64_t row_offset, FrGallery** out) {
    return guarded([&] {
        if (!out) throw ArgError{"out is null"};
        if (n > 0 && !rows_dev) throw ArgError{"rows is null"};
        FrGallery* g = new_gallery(n, dim, device, row_offset);
        try {
            if (n > 0) {
                FRB_CUDA(cudaMemcpyAsync(g->rows_f32, rows_dev, sizeof(float) * n * kDim, cudaMemcpyDeviceToDevice, g->stream));
                finish_rows(g, true);
            }
        } catch (...) {
            fr_gallery_destroy(g);
            throw;
        }
        *out = g;
    });
}

int fr_gallery_create_synthetic(int64_t n, int dim, uint64_t seed, int device, int64_t row_offset, FrGallery** out) {
    return guarded([&] {
        if (!out) throw ArgError{"out is null"};
        FrGallery* g = new_gallery(n, dim, device, row_offset);
        try {
            if (n > 0) {
                const int blocks = static_cast<int>(std::min<int64_t>((n + 7) / 8, g->sms * 16LL));
                synth_rows_kernel<<<blocks, 256, 0, g->stream>>>(g->rows_f32, g->rows_f16, n, seed, row_offset);
                count_launch();
                FRB_CUDA(cudaGetLastError());
                finish_rows(g, false);
            }
        } catch (...) {
            fr_gallery_destroy(g);
            throw;
        }
        *out = g;
    });
}

void fr_gallery_destroy(FrGallery* g) {
    if (!g) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    cudaFree(g->rows_f32);
    cudaFree(g->rows_f16);
    cudaFree(g->rows_f8);
    cudaFree(g->gmax);
    cudaFree(g->g4max);
    cudaFree(g->amax);
    cudaFree(g->w4max);
    cudaFree(g->q_gap);
    cudaFree(g->flagged_acc);
    cudaFree(g->q_dev);
    cudaFree(g->q_img);
    cudaFree(g->q_margin);
    cudaFree(g->cand_s);
    cudaFree(g->cand_i);
    cudaFree(g->flags);
    cudaFree(g->scan_ticket);
    cudaFree(g->gbest);
    cudaFree(g->app_buf);
    cudaFree(g->app_cnt);
    cudaFree(g->part_s);
    cudaFree(g->part_i);
    cudaFree(g->res_s);
    cudaFree(g->res_i);
    cudaFree(g->loc_s);
    cudaFree(g->loc_i);
    cudaFree(g->sims_ws);
    for (auto& e : g->ev_pool) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    for (auto& pr : g->fixed_ev)
        for (auto& e : pr)
            if (e) cudaEventDestroy(e);
    if (g->stream) cudaStreamDestroy(g->stream);
    if (prev >= 0) cudaSetDevice(prev);
    delete g;
}

int64_t fr_gallery_rows(const FrGallery* g) { return g ? g->n : -1; }
int fr_gallery_device(const FrGallery* g) { return g ? g->device : FR_EINVAL; }
int64_t fr_gallery_row_offset(const FrGallery* g) { return g ? g->row_offset : -1; }

int fr_gallery_set_path(FrGallery* g, int path) {
    return guarded([&] {
        if (!g) throw ArgError{"null gallery"};
        if (path != FR_PATH_AUTO && path != FR_PATH_EXACT && path != FR_PATH_TENSOR) throw ArgError{"unknown path"};
        g->path = path;
    });
}

int fr_gallery_set_scan(FrGallery* g, int scan) {
    return guarded([&] {
        if (!g) throw ArgError{"null gallery"};
        if (scan != FR_SCAN_F16 && scan != FR_SCAN_F8) throw ArgError{"unknown scan precision"};
        DeviceGuard dg(g->device);
        if (scan == FR_SCAN_F8 && !g->rows_f8 && g->capacity > 0) {
            float gmax = 0.f;
            FRB_CUDA(cudaMemcpy(&gmax, g->gmax, sizeof(float), cudaMemcpyDeviceToHost));
            if (gmax > 1.001f) throw StateError{"FR_SCAN_F8 needs L2-normalised rows (largest row norm > 1)"};
            // (a block-tiled copy, one contiguous 16 KiB chunk per TMA box, was measured against this row-major one on B200: no
            //  difference, 1.010 vs 1.015 ms per 10 M-row scan; the plain matrix stays)
            FRB_CUDA(cudaMalloc(&g->rows_f8, static_cast<size_t>(g->capacity) * kDim));
            if (g->n > 0) {
                const int blocks = static_cast<int>(std::min<int64_t>((g->n + 7) / 8, g->sms * 16LL));
                make_f8_copy_kernel<<<blocks, 256, 0, g->stream>>>(g->rows_f32, g->rows_f8, g->n, g->f8_seed, g->row_offset, g->g4max, g->w4max);
                count_launch();
                FRB_CUDA(cudaGetLastError());
                FRB_CUDA(cudaStreamSynchronize(g->stream));
                g->tmap8 = make_tmap_2d_u8(g->rows_f8, static_cast<uint64_t>(g->n), kDim, 128, 128);
            }
        }
        g->scan = scan;
    });
}

int64_t fr_gallery_capacity(const FrGallery* g) { return g ? g->capacity : -1; }

int fr_gallery_reserve(FrGallery* g, int64_t capacity) {
    return guarded([&] {
        if (!g) throw ArgError{"null gallery"};
        if (capacity < 0) throw ArgError{"negative capacity"};
        DeviceGuard dg(g->device);
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        grow_rows(g, capacity);
    });
}

int fr_gallery_append(FrGallery* g, const float* rows, int64_t n) {
    return guarded([&] {
        if (!g) throw ArgError{"null gallery"};
        if (n < 0 || (n > 0 && !rows)) throw ArgError{"bad rows / n"};
        if (n == 0) return;
        DeviceGuard dg(g->device);
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        if (g->rows_f8 || g->scan == FR_SCAN_F8) check_unit_rows(rows, n);
        if (g->n + n > g->capacity) grow_rows(g, std::max<int64_t>(g->n + n, g->capacity + g->capacity / 2 + 1024));
        float* dst = g->rows_f32 + g->n * kDim;
        FRB_CUDA(cudaMemcpyAsync(dst, rows, sizeof(float) * n * kDim, cudaMemcpyHostToDevice, g->stream));
        const int blocks = static_cast<int>(std::min<int64_t>((n + 7) / 8, g->sms * 16LL));
        // scan copies of the new rows only; gmax / g4max are running maxima (atomicMax), so they stay upper bounds
        make_scan_copy_kernel<<<blocks, 256, 0, g->stream>>>(dst, g->rows_f16 + g->n * kDim, n, g->gmax, g->amax);
        count_launch();
        if (g->rows_f8) {
            make_f8_copy_kernel<<<blocks, 256, 0, g->stream>>>(dst, g->rows_f8 + g->n * kDim, n, g->f8_seed, g->row_offset + g->n, g->g4max, g->w4max);
            count_launch();
        }
        FRB_CUDA(cudaGetLastError());
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        g->n += n;
        refresh_f16_ok(g);
        refresh_tmaps(g);
    });
}

int fr_gallery_update(FrGallery* g, int64_t first, const float* rows, int64_t n) {
    return guarded([&] {
        if (!g) throw ArgError{"null gallery"};
        if (n < 0 || (n > 0 && !rows)) throw ArgError{"bad rows / n"};
        if (first < 0 || first + n > g->n) throw ArgError{"row range out of bounds"};
        if (n == 0) return;
        DeviceGuard dg(g->device);
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        if (g->rows_f8 || g->scan == FR_SCAN_F8) check_unit_rows(rows, n);
        float* dst = g->rows_f32 + first * kDim;
        FRB_CUDA(cudaMemcpyAsync(dst, rows, sizeof(float) * n * kDim, cudaMemcpyHostToDevice, g->stream));
        const int blocks = static_cast<int>(std::min<int64_t>((n + 7) / 8, g->sms * 16LL));
        // the bounds (gmax, g4max, w4max) are running maxima: they keep the replaced rows' contribution and stay upper bounds
        make_scan_copy_kernel<<<blocks, 256, 0, g->stream>>>(dst, g->rows_f16 + first * kDim, n, g->gmax, g->amax);
        count_launch();
        if (g->rows_f8) {
            make_f8_copy_kernel<<<blocks, 256, 0, g->stream>>>(dst, g->rows_f8 + first * kDim, n, g->f8_seed, g->row_offset + first, g->g4max, g->w4max);
            count_launch();
        }
        FRB_CUDA(cudaGetLastError());
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        refresh_f16_ok(g);
    });
}

int fr_gallery_remove(FrGallery* g, int64_t row, int64_t* moved_from) {
    return guarded([&] {
        if (!g) throw ArgError{"null gallery"};
        if (row < 0 || row >= g->n) throw ArgError{"row out of range"};
        DeviceGuard dg(g->device);
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        const int64_t last = g->n - 1;
        if (row != last) {
            FRB_CUDA(cudaMemcpyAsync(g->rows_f32 + row * kDim, g->rows_f32 + last * kDim, sizeof(float) * kDim, cudaMemcpyDeviceToDevice, g->stream));
            FRB_CUDA(cudaMemcpyAsync(g->rows_f16 + row * kDim, g->rows_f16 + last * kDim, sizeof(__half) * kDim, cudaMemcpyDeviceToDevice, g->stream));
            if (g->rows_f8)
                FRB_CUDA(cudaMemcpyAsync(g->rows_f8 + row * kDim, g->rows_f8 + last * kDim, kDim, cudaMemcpyDeviceToDevice, g->stream));
            FRB_CUDA(cudaStreamSynchronize(g->stream));
        }
        if (moved_from) *moved_from = last;
        g->n = last;  // gmax / g4max keep the removed row's contribution: still upper bounds, the margins only get wider
        refresh_tmaps(g);
    });
}

int fr_gallery_clear(FrGallery* g) {
    return guarded([&] {
        if (!g) throw ArgError{"null gallery"};
        DeviceGuard dg(g->device);
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        g->n = 0;
        FRB_CUDA(cudaMemsetAsync(g->gmax, 0, sizeof(float), g->stream));
        FRB_CUDA(cudaMemsetAsync(g->g4max, 0, sizeof(float), g->stream));
        FRB_CUDA(cudaMemsetAsync(g->w4max, 0, sizeof(float), g->stream));
        FRB_CUDA(cudaMemsetAsync(g->amax, 0, sizeof(float), g->stream));
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        g->f16_ok = true;
    });
}

int fr_gallery_read_rows(FrGallery* g, int64_t first, int64_t count, float* out_rows) {
    return guarded([&] {
        if (!g || !out_rows) throw ArgError{"null argument"};
        if (first < 0 || count < 0 || first + count > g->n) throw ArgError{"row range out of bounds"};
        DeviceGuard dg(g->device);
        if (count == 0) return;
        FRB_CUDA(cudaMemcpyAsync(out_rows, g->rows_f32 + first * kDim, sizeof(float) * count * kDim, cudaMemcpyDeviceToHost, g->stream));
        FRB_CUDA(cudaStreamSynchronize(g->stream));
    });
}

int fr_gallery_sims_dev(FrGallery* g, const float* q_dev, int nq, float* out_dev, void* stream) {
    return guarded([&] {
        check_query_args(g, q_dev, nq);
        if (!out_dev) throw ArgError{"out is null"};
        DeviceGuard dg(g->device);
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g->stream;
        launch_sims(g, q_dev, nq, out_dev, st);
    });
}

int fr_gallery_sims(FrGallery* g, const float* q, int nq, float* out) {
    return guarded([&] {
        check_query_args(g, q, nq);
        if (!out) throw ArgError{"out is null"};
        DeviceGuard dg(g->device);
        // chunk the queries so that the device workspace stays bounded (the reference allocates n x m at once, src/matmul.cpp:41)
        const int64_t max_ws_floats = int64_t(1) << 28;  // 1 GiB
        const int step = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min(nq, kChunkQ), max_ws_floats / g->n)));
        for (int q0 = 0; q0 < nq; q0 += step) {
            const int m = std::min(step, nq - q0);
            ensure_sims_ws(g, static_cast<size_t>(m) * g->n);
            FRB_CUDA(cudaMemcpyAsync(g->q_dev, q + static_cast<size_t>(q0) * kDim, sizeof(float) * m * kDim, cudaMemcpyHostToDevice, g->stream));
            launch_sims(g, g->q_dev, m, g->sims_ws, g->stream);
            FRB_CUDA(cudaMemcpyAsync(out + static_cast<size_t>(q0) * g->n, g->sims_ws, sizeof(float) * m * g->n, cudaMemcpyDeviceToHost,
                                     g->stream));
            FRB_CUDA(cudaStreamSynchronize(g->stream));
        }
    });
}

int fr_gallery_topk_dev(FrGallery* g, const float* q_dev, int nq, int k, float* scores_dev, int64_t* idx_dev, void* stream) {
    return guarded([&] {
        check_query_args(g, q_dev, nq);
        if (k < 1 || k > FR_TOPK_MAX) throw ArgError{"k out of range"};
        if (!scores_dev || !idx_dev) throw ArgError{"null output"};
        DeviceGuard dg(g->device);
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g->stream;
        FrSearchStats total{};
        for (int q0 = 0; q0 < nq; q0 += kChunkQ) {
            const int m = std::min(kChunkQ, nq - q0);
            g->first_chunk = q0 == 0;
            topk_chunk(g, q_dev + static_cast<size_t>(q0) * kDim, m, k, scores_dev + static_cast<size_t>(q0) * k,
                       reinterpret_cast<long long*>(idx_dev) + static_cast<size_t>(q0) * k, st);
            add_stats(total, g->stats);
        }
        g->stats = total;
    });
}

int fr_gallery_topk(FrGallery* g, const float* q, int nq, int k, float* scores, int64_t* idx) {
    return guarded([&] {
        check_query_args(g, q, nq);
        if (k < 1 || k > FR_TOPK_MAX) throw ArgError{"k out of range"};
        if (!scores || !idx) throw ArgError{"null output"};
        DeviceGuard dg(g->device);
        FrSearchStats total{};
        for (int q0 = 0; q0 < nq; q0 += kChunkQ) {
            const int m = std::min(kChunkQ, nq - q0);
            FRB_CUDA(cudaMemcpyAsync(g->q_dev, q + static_cast<size_t>(q0) * kDim, sizeof(float) * m * kDim, cudaMemcpyHostToDevice, g->stream));
            g->first_chunk = q0 == 0;
            topk_chunk(g, g->q_dev, m, k, g->res_s, g->res_i, g->stream);
            FRB_CUDA(cudaMemcpyAsync(scores + static_cast<size_t>(q0) * k, g->res_s, sizeof(float) * m * k, cudaMemcpyDeviceToHost, g->stream));
            FRB_CUDA(cudaMemcpyAsync(idx + static_cast<size_t>(q0) * k, g->res_i, sizeof(long long) * m * k, cudaMemcpyDeviceToHost, g->stream));
            FRB_CUDA(cudaStreamSynchronize(g->stream));
            add_stats(total, g->stats);
        }
        g->stats = total;
    });
}

int fr_topk_merge_dev(const float* scores_parts_dev, const int64_t* idx_parts_dev, int n_parts, int nq, int k, float* scores_dev,
                      int64_t* idx_dev, int device, void* stream) {
    return guarded([&] {
        if (!scores_parts_dev || !idx_parts_dev || !scores_dev || !idx_dev) throw ArgError{"null argument"};
        if (n_parts < 1 || n_parts > 64 || nq < 1 || k < 1 || k > FR_TOPK_MAX) throw ArgError{"bad merge shape"};
        DeviceGuard dg(device);
        topk_merge_kernel<<<nq, 64, 0, static_cast<cudaStream_t>(stream)>>>(scores_parts_dev, reinterpret_cast<const long long*>(idx_parts_dev),
                                                                          n_parts, nq, k, scores_dev, reinterpret_cast<long long*>(idx_dev));
        count_launch();
        FRB_CUDA(cudaGetLastError());
    });
}

int fr_gallery_set_timing(FrGallery* g, int enable) {
    return guarded([&] {
        if (!g) throw ArgError{"null gallery"};
        g->timing = enable == 1;
        g->timing_fixed = enable == 2;
        g->ev_used = 0;
    });
}

int fr_gallery_pool_time(FrGallery* g, int count, double* total_ms) {
    return guarded([&] {
        if (!g || !total_ms) throw ArgError{"null argument"};
        if (count < 0 || static_cast<size_t>(count) > g->ev_pool.size()) throw ArgError{"more launches than event pairs in the pool"};
        DeviceGuard dg(g->device);
        double sum = 0;
        for (int i = 0; i < count; ++i) {
            FRB_CUDA(cudaEventSynchronize(g->ev_pool[i].second));
            float ms = 0;
            FRB_CUDA(cudaEventElapsedTime(&ms, g->ev_pool[i].first, g->ev_pool[i].second));
            sum += ms;
        }
        *total_ms = sum;
    });
}

int fr_gallery_last_scan_ms(FrGallery* g, int scan, double* ms) {
    return guarded([&] {
        if (!g || !ms) throw ArgError{"null argument"};
        if (scan != FR_SCAN_F16 && scan != FR_SCAN_F8) throw ArgError{"unknown scan precision"};
        if (!g->fixed_ev[scan][0] || !g->fixed_ev[scan][1]) throw StateError{"no search ran in timing mode 2 on this scan copy"};
        DeviceGuard dg(g->device);
        FRB_CUDA(cudaEventSynchronize(g->fixed_ev[scan][1]));
        float t = 0;
        FRB_CUDA(cudaEventElapsedTime(&t, g->fixed_ev[scan][0], g->fixed_ev[scan][1]));
        *ms = t;
    });
}

int fr_gallery_scan_time(FrGallery* g, double* total_ms, int* launches) {
    return guarded([&] {
        if (!g || !total_ms || !launches) throw ArgError{"null argument"};
        DeviceGuard dg(g->device);
        double sum = 0;
        for (size_t i = 0; i < g->ev_used; ++i) {
            FRB_CUDA(cudaEventSynchronize(g->ev_pool[i].second));
            float ms = 0;
            FRB_CUDA(cudaEventElapsedTime(&ms, g->ev_pool[i].first, g->ev_pool[i].second));
            sum += ms;
        }
        *total_ms = sum;
        *launches = static_cast<int>(g->ev_used);
        g->ev_used = 0;
    });
}

int fr_gallery_last_flagged(FrGallery* g, int* out) {
    return guarded([&] {
        if (!g || !out) throw ArgError{"null argument"};
        DeviceGuard dg(g->device);
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        FRB_CUDA(cudaMemcpy(out, g->flagged_acc, sizeof(int), cudaMemcpyDeviceToHost));  // summed over the call's 256-query chunks
    });
}

int fr_gallery_debug_read(FrGallery* g, int what, int64_t first, int64_t count, void* out) {
    return guarded([&] {
        if (!g || !out) throw ArgError{"null argument"};
        DeviceGuard dg(g->device);
        FRB_CUDA(cudaStreamSynchronize(g->stream));
        switch (what) {
        case 0:
            if (!g->rows_f8) throw StateError{"no e4m3 scan copy (fr_gallery_set_scan(FR_SCAN_F8) first)"};
            if (first < 0 || count < 0 || first + count > g->n) throw ArgError{"row range out of bounds"};
            FRB_CUDA(cudaMemcpy(out, g->rows_f8 + first * kDim, static_cast<size_t>(count) * kDim, cudaMemcpyDeviceToHost));
            break;
        case 1:
            FRB_CUDA(cudaMemcpy(out, g->q_img, static_cast<size_t>(kChunkQ) * kDim * (g->scan == FR_SCAN_F8 ? 1 : 2), cudaMemcpyDeviceToHost));
            break;
        case 2: FRB_CUDA(cudaMemcpy(out, g->q_margin, sizeof(float) * kChunkQ, cudaMemcpyDeviceToHost)); break;
        case 3: FRB_CUDA(cudaMemcpy(out, g->q_gap, sizeof(float) * kChunkQ, cudaMemcpyDeviceToHost)); break;
        case 4: {
            float* o = static_cast<float*>(out);
            FRB_CUDA(cudaMemcpy(o, g->gmax, sizeof(float), cudaMemcpyDeviceToHost));
            FRB_CUDA(cudaMemcpy(o + 1, g->g4max, sizeof(float), cudaMemcpyDeviceToHost));
            FRB_CUDA(cudaMemcpy(o + 2, g->w4max, sizeof(float), cudaMemcpyDeviceToHost));
            break;
        }
        case 5: FRB_CUDA(cudaMemcpy(out, g->cand_s, sizeof(float) * kMaxLists * kChunkQ * 16, cudaMemcpyDeviceToHost)); break;
        case 6: FRB_CUDA(cudaMemcpy(out, g->cand_i, sizeof(int) * kMaxLists * kChunkQ * 16, cudaMemcpyDeviceToHost)); break;
        default: throw ArgError{"unknown debug read"};
        }
    });
}

int fr_gallery_last_stats(const FrGallery* g, FrSearchStats* out) {
    return guarded([&] {
        if (!g || !out) throw ArgError{"null argument"};
        *out = g->stats;
    });
}

}  // extern "C"

#include "exchange_impl.cuh"
