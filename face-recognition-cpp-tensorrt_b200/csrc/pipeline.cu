// End-to-end pipeline on one GPU: detect -> face-list compaction -> crop + bicubic resize -> embed -> search, all on the device.
// The frames travel over PCIe in sub-batches on a copy stream while earlier sub-batches are already being detected; the face
// list is compacted by a kernel (no host round trip between detector and embedder): the host only learns the face COUNT, while the
// first embedder chunk is already running. Replaces the body of the reference's /inference handler (/root/reference src/app.cpp:293-352):
// detector.findFace (src/retinaface.cpp:147-152) -> recognizer.forward (src/arcface.cpp:166-187: getCroppedFaces :3-17,
// preprocessFaces :116-129, doInference :139-148) -> featureMatching + getOutputs (:189-217).
#include <algorithm>
#include <memory>
#include <vector>

#include "common.h"

// internal hooks (detector.cu / embedder.cu)
namespace frb {
void detector_forward_dev(FrDetector* d, const uint8_t* frames_dev, int stride, int batch, cudaStream_t st, int slot0);
const FrBbox* detector_boxes(const FrDetector* d);
const int* detector_counts(const FrDetector* d);
void detector_dims(const FrDetector* d, int* frame_h, int* frame_w, int* max_batch, int* max_faces, int* device);
uint8_t* embedder_u8_input(FrEmbedder* e);
float* embedder_output(FrEmbedder* e);
int embedder_max_batch(const FrEmbedder* e);
int embedder_device(const FrEmbedder* e);
void embedder_forward_u8(FrEmbedder* e, int batch, cudaStream_t st);
cudaStream_t embedder_stream(FrEmbedder* e);
uint8_t* embedder_frame_scratch(FrEmbedder* e, size_t bytes);
void* embedder_faces_scratch(FrEmbedder* e, size_t bytes);
}  // namespace frb

using namespace frb;

namespace {

struct FaceRef {
    int frame;
    int x1, y1, x2, y2;  // Bbox semantics: x = row, y = column
};

// cubic convolution coefficients, A = -0.75 (OpenCV interpolateCubic, imgproc/src/resize.cpp), float arithmetic without contraction
__device__ __forceinline__ void cubic_coeffs(float x, float (&c)[4]) {
    const float A = -0.75f;
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, __fadd_rn(x, 1.f)), __fmul_rn(5.f, A)), __fadd_rn(x, 1.f)), __fmul_rn(8.f, A)),
                               __fadd_rn(x, 1.f)),
                     __fmul_rn(4.f, A));
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), x), __fadd_rn(A, 3.f)), x), x), 1.f);
    const float y = __fsub_rn(1.f, x);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), y), __fadd_rn(A, 3.f)), y), y), 1.f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

// One axis of OpenCV's resize set-up (modules/imgproc/src/resize.cpp, hal::resize + resizeGeneric_ tables): destination index d of a
// src -> 112 resize: scale = 1 / ((double)112 / src) (NOT src / 112: OpenCV inverts inv_scale), f = (float)((d + 0.5) * scale - 0.5),
// s = floor(f), f -= s, float cubic coefficients, then the fixed-point table  saturate_cast<short>(c * INTER_RESIZE_COEF_SCALE)
// with INTER_RESIZE_COEF_BITS = 11 (round half to even). No fused multiply-adds: the x86 baseline build has none.
__device__ __forceinline__ void cubic_axis(int d, int src, int& s, int (&ic)[4]) {
    const double inv_scale = __ddiv_rn(112.0, static_cast<double>(src));
    const double scale = __ddiv_rn(1.0, inv_scale);
    float f = static_cast<float>(__dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), 0.5));
    s = static_cast<int>(floorf(f));
    f = __fsub_rn(f, static_cast<float>(s));
    float c[4];
    cubic_coeffs(f, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) ic[k] = __float2int_rn(__fmul_rn(c[k], 2048.f));
}

// getCroppedFaces (src/arcface.cpp:3-17): ROI = Rect(Point(y1, x1), Point(y2, x2)) = rows [x1, x2), columns [y1, y2) of the frame,
// cv::resize(..., Size(112, 112), INTER_CUBIC). Byte-exact restatement of OpenCV's own u8 bicubic (resizeGeneric_<HResizeCubic<uchar,
// int, short>, VResizeCubic<..., VResizeCubicVec_32s8u>>): horizontal pass in integers with the 11-bit coefficient table (taps
// clamped to the ROI), vertical pass as the SIMD kernel computes it — float, beta * 2^-22, accumulated from tap 3 down to tap 0 with
// separate multiplies and adds, round half to even, saturate. (OpenCV is third-party code the reference links, README.md:11; the
// oracle is cv2 with its closed-source IPP accelerator switched off, tests/test_pipeline_gpu.py.)
// One thread per output pixel; output u8 BGR HWC = CroppedFace::face, and the embedder's input.
__global__ void __launch_bounds__(256) crop_resize_kernel(const uint8_t* __restrict__ frames, int frame_h, int frame_w, int stride,
                                                          const FaceRef* __restrict__ faces, const int* __restrict__ n_faces_dev, int first,
                                                          int n_faces, uint8_t* __restrict__ out) {
    constexpr int D = 112;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int face = blockIdx.y;  // faces[face] is face `first + face` of the batch
    // device-side face list (fr_pipeline_run): the total count lives on the device; chunk slots beyond it are skipped
    if (n_faces_dev) n_faces = min(n_faces, *n_faces_dev - first);
    if (t >= D * D || face >= n_faces) return;
    const int dy = t / D, dx = t % D;
    const FaceRef f = faces[face];
    const int r0 = min(f.x1, f.x2), c0 = min(f.y1, f.y2);
    const int sh = max(abs(f.x2 - f.x1), 1), sw = max(abs(f.y2 - f.y1), 1);  // an empty ROI would throw in OpenCV; use 1 pixel
    const uint8_t* base = frames + static_cast<size_t>(f.frame) * frame_h * stride;
    int sx, sy, ia[4], ib[4];
    cubic_axis(dx, sw, sx, ia);
    cubic_axis(dy, sh, sy, ib);
    int cols[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) cols[i] = min(max(min(max(sx - 1 + i, 0), sw - 1) + c0, 0), frame_w - 1) * 3;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 3; j >= 0; --j) {
        const int rr = min(max(sy - 1 + j, 0), sh - 1) + r0;
        const uint8_t* row = base + static_cast<size_t>(min(max(rr, 0), frame_h - 1)) * stride;
        const float b = __fmul_rn(static_cast<float>(ib[j]), 1.f / 4194304.f);  // beta * 1 / (INTER_RESIZE_COEF_SCALE^2): exact
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            int h = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) h += static_cast<int>(row[cols[i] + ch]) * ia[i];
            const float p = __fmul_rn(static_cast<float>(h), b);
            acc[ch] = j == 3 ? p : __fadd_rn(p, acc[ch]);
        }
    }
    uint8_t* o = out + (static_cast<size_t>(face) * D * D + t) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) o[ch] = static_cast<uint8_t>(min(max(__float2int_rn(acc[ch]), 0), 255));
}

// Face list on the device: face k of the batch = (frame, box) in frame-major, detection order — the order the host loop of the
// reference visits them (src/app.cpp:304-310). One block; exclusive scan of the per-frame counts.
__global__ void __launch_bounds__(256) compact_faces_kernel(const FrBbox* __restrict__ boxes, const int* __restrict__ counts, int batch,
                                                            int max_faces, FaceRef* __restrict__ faces, int* __restrict__ n_out) {
    __shared__ int warp_tot[8];
    __shared__ int base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int f0 = 0; f0 < batch; f0 += blockDim.x) {
        const int f = f0 + threadIdx.x;
        const int c = f < batch ? min(max(counts[f], 0), max_faces) : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        off += inc - c;
        for (int j = 0; j < c; ++j) {
            const FrBbox b = boxes[static_cast<size_t>(f) * max_faces + j];
            faces[off + j] = FaceRef{f, b.x1, b.y1, b.x2, b.y2};
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += warp_tot[w];
            base_s += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = base_s;
}

}  // namespace

// One batch in flight. The pipeline keeps kSlots of them so that a batch's H2D copy and host-side bookkeeping overlap the previous
// batch's kernels (fr_pipeline_submit / fr_pipeline_collect); the detector's and the embedder's own buffers exist once - kernels of
// consecutive batches simply queue on the one compute stream.
struct PipeSlot {
    uint8_t* frames_dev = nullptr;   // max_batch x frame_h x frame_w x 3 (the crop kernels read it after the detector did)
    FaceRef* faces_dev = nullptr;
    int* n_dev = nullptr;
    float* embeds_dev = nullptr;     // max_batch * max_faces x 512
    float* score_dev = nullptr;
    long long* idx_dev = nullptr;
    int* h_n = nullptr;              // pinned
    FrBbox* h_boxes = nullptr;       // pinned, max_batch * max_faces
    int* h_counts = nullptr;         // pinned, max_batch
    float* h_score = nullptr;        // pinned
    long long* h_idx = nullptr;      // pinned
    float* h_embeds = nullptr;       // pinned, allocated on first use
    std::vector<cudaEvent_t> h2d_done;
    cudaEvent_t count_ready = nullptr, done = nullptr;
    int batch = 0, embedded = 0;     // frames of the batch in flight; faces covered by the chunks enqueued so far
    int stride = 0;
    bool want_embeds = false, busy = false;
};

constexpr int kSlots = 2;

struct FrPipeline {
    FrDetector* det = nullptr;
    FrEmbedder* emb = nullptr;
    FrGallery* gal = nullptr;
    int device = 0, frame_h = 0, frame_w = 0, max_batch = 0, max_faces = 0, emb_batch = 0;
    int sub_batch = 16;              // frames per H2D / detect sub-batch (FR_PIPE_SUB)
    int spec = 0;                    // faces embedded before the host knows the count (adapts to the previous batch's face count)
    cudaStream_t stream = nullptr;   // compute
    cudaStream_t copy_stream = nullptr;
    PipeSlot slot[kSlots];
    int head = 0, tail = 0, in_flight = 0;  // ring of submitted batches: collect takes `tail`, submit fills `head`
};

namespace {

void crop_into_embedder(FrPipeline* p, const uint8_t* frames_dev, int stride, const FaceRef* faces_dev, const int* n_dev, int beg, int m,
                        cudaStream_t st) {
    dim3 grid((112 * 112 + 255) / 256, m);
    // faces beyond the device-side count leave their (stale) crop untouched; their embeddings are never read
    crop_resize_kernel<<<grid, 256, 0, st>>>(frames_dev, p->frame_h, p->frame_w, stride, faces_dev + beg, n_dev, beg, m, embedder_u8_input(p->emb));
    count_launch();
    FRB_CUDA(cudaGetLastError());
}

void embed_chunk(FrPipeline* p, PipeSlot& s, int stride, int beg, int m, cudaStream_t st) {
    NvtxRange nvtx("fr.pipeline.crop_embed");
    crop_into_embedder(p, s.frames_dev, stride, s.faces_dev, s.n_dev, beg, m, st);
    embedder_forward_u8(p->emb, m, st);
    FRB_CUDA(cudaMemcpyAsync(s.embeds_dev + static_cast<size_t>(beg) * 512, embedder_output(p->emb), sizeof(float) * m * 512,
                             cudaMemcpyDeviceToDevice, st));
}

// search + result copies of the batch in `s` for its n faces (n known on the host, or an upper bound: see enqueue)
void enqueue_search(FrPipeline* p, PipeSlot& s, int n, cudaStream_t st) {
    if (n <= 0) return;
    if (s.want_embeds) FRB_CUDA(cudaMemcpyAsync(s.h_embeds, s.embeds_dev, sizeof(float) * n * 512, cudaMemcpyDeviceToHost, st));
    if (p->gal && fr_gallery_rows(p->gal) > 0) {
        const int rc = fr_gallery_topk_dev(p->gal, s.embeds_dev, n, 1, s.score_dev, reinterpret_cast<int64_t*>(s.idx_dev), st);
        if (rc != FR_OK) throw CudaError{std::string("pipeline search failed: ") + fr_last_error()};
        FRB_CUDA(cudaMemcpyAsync(s.h_score, s.score_dev, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
        FRB_CUDA(cudaMemcpyAsync(s.h_idx, s.idx_dev, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    }
}

// Enqueue one batch (host frames) into slot `s` without waiting for anything on the host: sub-batch H2D on the copy stream, detection,
// face-list compaction, and the embedder + search for the first `spec` faces (a guess: the previous batch's count). finish() completes
// the rare batch that has more faces than the guess.
void enqueue(FrPipeline* p, PipeSlot& s, const uint8_t* frames, int stride, int batch, bool want_embeds) {
    NvtxRange nvtx("fr.pipeline.submit");
    cudaStream_t st = p->stream, cs = p->copy_stream;
    const size_t row_bytes = static_cast<size_t>(p->frame_w) * 3, frame_bytes = row_bytes * p->frame_h;
    if (want_embeds && !s.h_embeds) FRB_CUDA(cudaMallocHost(&s.h_embeds, sizeof(float) * 512 * p->max_batch * p->max_faces));
    s.batch = batch;
    s.stride = static_cast<int>(row_bytes);
    s.want_embeds = want_embeds;
    // the slot's previous batch was collected (its kernels are done), so the copies may start at once
    // Sub-batches let the detector start while later frames still cross PCIe - worth it only when the GPU would otherwise wait for the
    // copy. With another batch in flight the copy is hidden behind that batch's kernels and the detector is more efficient on the whole
    // batch at once (measured: 2.15 ms for 64 frames in one go, 4 x 0.8 ms in sub-batches of 16).
    const int sub = p->in_flight > 0 ? batch : p->sub_batch;
    int k = 0;
    for (int f0 = 0; f0 < batch; f0 += sub, ++k) {
        const int m = std::min(sub, batch - f0);
        FRB_CUDA(cudaMemcpy2DAsync(s.frames_dev + f0 * frame_bytes, row_bytes, frames + static_cast<size_t>(f0) * p->frame_h * stride, stride,
                                   row_bytes, static_cast<size_t>(m) * p->frame_h, cudaMemcpyHostToDevice, cs));
        FRB_CUDA(cudaEventRecord(s.h2d_done[k], cs));
        FRB_CUDA(cudaStreamWaitEvent(st, s.h2d_done[k], 0));
        detector_forward_dev(p->det, s.frames_dev + f0 * frame_bytes, static_cast<int>(row_bytes), m, st, f0);
    }
    compact_faces_kernel<<<1, 256, 0, st>>>(detector_boxes(p->det), detector_counts(p->det), batch, p->max_faces, s.faces_dev, s.n_dev);
    count_launch();
    FRB_CUDA(cudaGetLastError());
    // the detector's result buffers are overwritten by the next batch's detection: copy them out now (stream order)
    FRB_CUDA(cudaMemcpyAsync(s.h_n, s.n_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaMemcpyAsync(s.h_boxes, detector_boxes(p->det), sizeof(FrBbox) * batch * p->max_faces, cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaMemcpyAsync(s.h_counts, detector_counts(p->det), sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaEventRecord(s.count_ready, st));
    // speculative chunks: enqueued before the host knows the face count, so the GPU never waits for the host in the middle
    const int slots = batch * p->max_faces;
    const int guess = std::min(slots, p->spec > 0 ? p->spec : slots);
    s.embedded = 0;
    for (int beg = 0; beg < guess; beg += p->emb_batch) {
        embed_chunk(p, s, s.stride, beg, std::min(p->emb_batch, guess - beg), st);
        s.embedded = std::min(guess, beg + p->emb_batch);
    }
    // the search of the guessed range; rows beyond the real count hold stale embeddings and are dropped by finish()
    enqueue_search(p, s, s.embedded, st);
    FRB_CUDA(cudaEventRecord(s.done, st));
    s.busy = true;
}

// Wait for the batch in `s`; returns its face count. Runs the remaining chunks first if the batch had more faces than were guessed.
int finish(FrPipeline* p, PipeSlot& s) {
    NvtxRange nvtx("fr.pipeline.collect");
    cudaStream_t st = p->stream;
    FRB_CUDA(cudaEventSynchronize(s.count_ready));
    const int n = *s.h_n;
    p->spec = std::max(std::min(32, p->emb_batch), (n + 31) / 32 * 32);
    if (n > s.embedded) {
        // more faces than guessed: embed the rest (chunked like ArcFaceIR50::forward, src/arcface.cpp:177-185) and search again.
        // The embedder's crop buffer may by now hold a LATER batch's crops - it is rewritten here, and that later batch's kernels have
        // already consumed it in stream order.
        for (int beg = s.embedded; beg < n; beg += p->emb_batch) embed_chunk(p, s, s.stride, beg, std::min(p->emb_batch, n - beg), st);
        s.embedded = n;
        enqueue_search(p, s, n, st);
        FRB_CUDA(cudaEventRecord(s.done, st));
    }
    FRB_CUDA(cudaEventSynchronize(s.done));
    if (!(p->gal && fr_gallery_rows(p->gal) > 0)) {
        std::fill(s.h_idx, s.h_idx + n, -1LL);
        std::fill(s.h_score, s.h_score + n, 0.f);
    }
    s.busy = false;
    return n;
}

// per-frame face slots: face j of frame f -> slot f * max_faces + j; the device results are compact (face order)
void scatter_results(FrPipeline* p, PipeSlot& s, int n, FrBbox* boxes, int* counts, int64_t* top1_idx, float* top1_score, float* embeddings) {
    const int batch = s.batch;
    std::copy(s.h_boxes, s.h_boxes + static_cast<size_t>(batch) * p->max_faces, boxes);
    std::copy(s.h_counts, s.h_counts + batch, counts);
    int k = 0;
    for (int f = 0; f < batch; ++f)
        for (int j = 0; j < p->max_faces; ++j) {
            const size_t slot = static_cast<size_t>(f) * p->max_faces + j;
            const bool live = j < counts[f];
            if (top1_idx) top1_idx[slot] = live ? s.h_idx[k] : -1;
            if (top1_score) top1_score[slot] = live ? s.h_score[k] : 0.f;
            if (embeddings) {
                if (live) std::copy(s.h_embeds + static_cast<size_t>(k) * 512, s.h_embeds + static_cast<size_t>(k + 1) * 512, embeddings + slot * 512);
                else std::fill(embeddings + slot * 512, embeddings + (slot + 1) * 512, 0.f);
            }
            if (live) ++k;
        }
    if (k != n) throw CudaError{"pipeline: device face count disagrees with the per-frame counts"};
}

void check_submit(const FrPipeline* p, const uint8_t* frames, int stride, int batch) {
    if (!p || !frames) throw ArgError{"null argument"};
    if (batch < 1 || batch > p->max_batch) throw ArgError{"batch out of range (1..max_batch)"};
    if (stride < p->frame_w * 3) throw ArgError{"stride smaller than frame_w * 3"};
}

void drain(FrPipeline* p) {  // after an error: leave no work in flight that still reads the caller's buffers
    cudaStreamSynchronize(p->copy_stream);
    cudaStreamSynchronize(p->stream);
    for (auto& s : p->slot) s.busy = false;
    p->head = p->tail = p->in_flight = 0;
}

}  // namespace

// internal hook for the request batcher (service.cu)
namespace frb {
void pipeline_dims(const FrPipeline* p, int* device, int* frame_h, int* frame_w, int* max_batch, int* max_faces) {
    *device = p->device;
    *frame_h = p->frame_h;
    *frame_w = p->frame_w;
    *max_batch = p->max_batch;
    *max_faces = p->max_faces;
}
}  // namespace frb

extern "C" {

int fr_pipeline_create(FrDetector* det, FrEmbedder* emb, FrGallery* gal, FrPipeline** out) {
    return guarded([&] {
        if (!det || !emb || !out) throw ArgError{"null argument"};
        std::unique_ptr<FrPipeline> p(new FrPipeline());
        p->det = det;
        p->emb = emb;
        p->gal = gal;
        detector_dims(det, &p->frame_h, &p->frame_w, &p->max_batch, &p->max_faces, &p->device);
        if (embedder_device(emb) != p->device) throw ArgError{"detector and embedder live on different devices"};
        if (gal && fr_gallery_device(gal) != p->device) throw ArgError{"the gallery lives on a different device than the detector"};
        p->emb_batch = embedder_max_batch(emb);
        if (const char* e = std::getenv("FR_PIPE_SUB")) p->sub_batch = std::max(1, std::atoi(e));
        DeviceGuard dg(p->device);
        const size_t slots = static_cast<size_t>(p->max_batch) * p->max_faces;
        const size_t frame_bytes = static_cast<size_t>(p->frame_h) * p->frame_w * 3;
        try {
            FRB_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
            FRB_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
            for (PipeSlot& s : p->slot) {
                s.h2d_done.resize((p->max_batch + p->sub_batch - 1) / p->sub_batch);
                for (auto& e : s.h2d_done) FRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                FRB_CUDA(cudaEventCreateWithFlags(&s.count_ready, cudaEventDisableTiming));
                FRB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
                FRB_CUDA(cudaMalloc(&s.frames_dev, frame_bytes * p->max_batch));
                FRB_CUDA(cudaMalloc(&s.faces_dev, sizeof(FaceRef) * slots));
                FRB_CUDA(cudaMalloc(&s.n_dev, sizeof(int)));
                FRB_CUDA(cudaMalloc(&s.embeds_dev, sizeof(float) * slots * 512));
                FRB_CUDA(cudaMalloc(&s.score_dev, sizeof(float) * slots));
                FRB_CUDA(cudaMalloc(&s.idx_dev, sizeof(long long) * slots));
                FRB_CUDA(cudaMallocHost(&s.h_n, sizeof(int)));
                FRB_CUDA(cudaMallocHost(&s.h_boxes, sizeof(FrBbox) * slots));
                FRB_CUDA(cudaMallocHost(&s.h_counts, sizeof(int) * p->max_batch));
                FRB_CUDA(cudaMallocHost(&s.h_score, sizeof(float) * slots));
                FRB_CUDA(cudaMallocHost(&s.h_idx, sizeof(long long) * slots));
            }
        } catch (...) {
            fr_pipeline_destroy(p.release());
            throw;
        }
        *out = p.release();
    });
}

void fr_pipeline_destroy(FrPipeline* p) {
    if (!p) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (PipeSlot& s : p->slot) {
        cudaFree(s.frames_dev);
        cudaFree(s.faces_dev);
        cudaFree(s.n_dev);
        cudaFree(s.embeds_dev);
        cudaFree(s.score_dev);
        cudaFree(s.idx_dev);
        cudaFreeHost(s.h_n);
        cudaFreeHost(s.h_boxes);
        cudaFreeHost(s.h_counts);
        cudaFreeHost(s.h_score);
        cudaFreeHost(s.h_idx);
        cudaFreeHost(s.h_embeds);
        for (auto e : s.h2d_done)
            if (e) cudaEventDestroy(e);
        if (s.count_ready) cudaEventDestroy(s.count_ready);
        if (s.done) cudaEventDestroy(s.done);
    }
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (prev >= 0) cudaSetDevice(prev);
    delete p;
}

int fr_pipeline_submit(FrPipeline* p, const uint8_t* frames, int stride, int batch, int want_embeddings) {
    return guarded([&] {
        check_submit(p, frames, stride, batch);
        if (p->in_flight >= kSlots) throw StateError{"fr_pipeline_submit: two batches are already in flight, collect one first"};
        DeviceGuard dg(p->device);
        try {
            enqueue(p, p->slot[p->head], frames, stride, batch, want_embeddings != 0);
        } catch (...) {
            drain(p);
            throw;
        }
        p->head = (p->head + 1) % kSlots;
        ++p->in_flight;
    });
}

int fr_pipeline_collect(FrPipeline* p, FrBbox* boxes, int* counts, int64_t* top1_idx, float* top1_score, float* embeddings) {
    return guarded([&] {
        if (!p || !boxes || !counts) throw ArgError{"null argument"};
        if (p->in_flight < 1) throw StateError{"fr_pipeline_collect: nothing was submitted"};
        PipeSlot& s = p->slot[p->tail];
        if (embeddings && !s.want_embeds) throw ArgError{"embeddings requested from a batch submitted without want_embeddings"};
        DeviceGuard dg(p->device);
        try {
            const int n = finish(p, s);
            scatter_results(p, s, n, boxes, counts, top1_idx, top1_score, embeddings);
        } catch (...) {
            drain(p);
            throw;
        }
        p->tail = (p->tail + 1) % kSlots;
        --p->in_flight;
    });
}

int fr_pipeline_in_flight(const FrPipeline* p) { return p ? p->in_flight : FR_EINVAL; }

int fr_pipeline_run(FrPipeline* p, const uint8_t* frames, int stride, int batch, FrBbox* boxes, int* counts, int64_t* top1_idx,
                    float* top1_score, float* embeddings) {
    return guarded([&] {
        check_submit(p, frames, stride, batch);
        if (!boxes || !counts) throw ArgError{"null argument"};
        if (p->in_flight != 0) throw StateError{"fr_pipeline_run: batches are in flight (collect them first)"};
        DeviceGuard dg(p->device);
        PipeSlot& s = p->slot[0];
        try {
            enqueue(p, s, frames, stride, batch, embeddings != nullptr);
            const int n = finish(p, s);
            scatter_results(p, s, n, boxes, counts, top1_idx, top1_score, embeddings);
        } catch (...) {
            drain(p);
            throw;
        }
    });
}

/* getCroppedFaces + preprocessFaces + doInference for caller-provided boxes of ONE frame (ArcFaceIR50::forward,
 * src/arcface.cpp:166-187). crops_u8 (optional): n x 112 x 112 x 3 BGR = CroppedFace::face. Host buffers. The frame and face-list
 * device buffers belong to the embedder and are reused from call to call (the reference's forward runs once per video frame). */
int fr_embedder_run_boxes(FrEmbedder* e, const uint8_t* frame, int frame_h, int frame_w, int stride, const FrBbox* boxes, int n, float* out512,
                          uint8_t* crops_u8) {
    return guarded([&] {
        if (!e || !frame || !boxes || !out512) throw ArgError{"null argument"};
        if (n < 1) throw ArgError{"no boxes"};
        if (frame_h < 1 || frame_w < 1 || stride < frame_w * 3) throw ArgError{"bad frame geometry"};
        DeviceGuard dg(embedder_device(e));
        cudaStream_t st = embedder_stream(e);
        FRB_CUDA(cudaStreamSynchronize(st));
        uint8_t* fdev = embedder_frame_scratch(e, static_cast<size_t>(frame_h) * frame_w * 3);
        FaceRef* faces_dev = static_cast<FaceRef*>(embedder_faces_scratch(e, sizeof(FaceRef) * std::max(n, 64)));
        FRB_CUDA(cudaMemcpy2DAsync(fdev, static_cast<size_t>(frame_w) * 3, frame, stride, static_cast<size_t>(frame_w) * 3, frame_h,
                                   cudaMemcpyHostToDevice, st));
        std::vector<FaceRef> faces(n);
        for (int i = 0; i < n; ++i) faces[i] = FaceRef{0, boxes[i].x1, boxes[i].y1, boxes[i].x2, boxes[i].y2};
        FRB_CUDA(cudaMemcpyAsync(faces_dev, faces.data(), sizeof(FaceRef) * n, cudaMemcpyHostToDevice, st));
        try {
            const int mb = embedder_max_batch(e);
            for (int beg = 0; beg < n; beg += mb) {
                const int m = std::min(mb, n - beg);
                dim3 grid((112 * 112 + 255) / 256, m);
                crop_resize_kernel<<<grid, 256, 0, st>>>(fdev, frame_h, frame_w, frame_w * 3, faces_dev + beg, nullptr, 0, m, embedder_u8_input(e));
                count_launch();
                if (crops_u8)
                    FRB_CUDA(cudaMemcpyAsync(crops_u8 + static_cast<size_t>(beg) * 112 * 112 * 3, embedder_u8_input(e),
                                             static_cast<size_t>(m) * 112 * 112 * 3, cudaMemcpyDeviceToHost, st));
                embedder_forward_u8(e, m, st);
                FRB_CUDA(cudaMemcpyAsync(out512 + static_cast<size_t>(beg) * 512, embedder_output(e), sizeof(float) * m * 512, cudaMemcpyDeviceToHost, st));
            }
            FRB_CUDA(cudaStreamSynchronize(st));
        } catch (...) {
            cudaStreamSynchronize(st);  // `faces` (pageable host memory) must outlive its copy
            throw;
        }
    });
}

}  // extern "C"
