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
uint8_t* detector_frames_buffer(FrDetector* d);
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

struct FrPipeline {
    FrDetector* det = nullptr;
    FrEmbedder* emb = nullptr;
    FrGallery* gal = nullptr;
    int device = 0, frame_h = 0, frame_w = 0, max_batch = 0, max_faces = 0, emb_batch = 0;
    int sub_batch = 16;              // frames per H2D / detect sub-batch (FR_PIPE_SUB)
    int spec = 0;                    // size of the speculative first embedder chunk (adapts to the previous call's face count)
    cudaStream_t stream = nullptr;   // compute
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> h2d_done;
    cudaEvent_t count_ready = nullptr;
    FaceRef* faces_dev = nullptr;
    int* n_dev = nullptr;
    int* h_n = nullptr;              // pinned
    float* embeds_dev = nullptr;     // max_batch * max_faces x 512
    float* score_dev = nullptr;
    long long* idx_dev = nullptr;
    FrBbox* h_boxes = nullptr;       // pinned, max_batch * max_faces
    int* h_counts = nullptr;         // pinned, max_batch
    float* h_score = nullptr;        // pinned
    long long* h_idx = nullptr;      // pinned
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

void embed_chunk(FrPipeline* p, const uint8_t* frames_dev, int stride, int beg, int m, cudaStream_t st) {
    NvtxRange nvtx("fr.pipeline.crop_embed");
    crop_into_embedder(p, frames_dev, stride, p->faces_dev, p->n_dev, beg, m, st);
    embedder_forward_u8(p->emb, m, st);
    FRB_CUDA(cudaMemcpyAsync(p->embeds_dev + static_cast<size_t>(beg) * 512, embedder_output(p->emb), sizeof(float) * m * 512,
                             cudaMemcpyDeviceToDevice, st));
}

// frames: host. Fills the pinned host arrays of the pipeline; returns the number of faces.
int run_batch(FrPipeline* p, const uint8_t* frames, int stride, int batch, float* embeddings_host) {
    NvtxRange nvtx("fr.pipeline.batch");
    cudaStream_t st = p->stream, cs = p->copy_stream;
    uint8_t* fdev = detector_frames_buffer(p->det);
    const size_t row_bytes = static_cast<size_t>(p->frame_w) * 3, frame_bytes = row_bytes * p->frame_h;
    // the previous call's kernels no longer read the frame buffer (every call ends with a stream sync), so the copies may start now
    int k = 0;
    for (int f0 = 0; f0 < batch; f0 += p->sub_batch, ++k) {
        const int m = std::min(p->sub_batch, batch - f0);
        FRB_CUDA(cudaMemcpy2DAsync(fdev + f0 * frame_bytes, row_bytes, frames + static_cast<size_t>(f0) * p->frame_h * stride, stride, row_bytes,
                                   static_cast<size_t>(m) * p->frame_h, cudaMemcpyHostToDevice, cs));
        FRB_CUDA(cudaEventRecord(p->h2d_done[k], cs));
        FRB_CUDA(cudaStreamWaitEvent(st, p->h2d_done[k], 0));
        detector_forward_dev(p->det, fdev + f0 * frame_bytes, static_cast<int>(row_bytes), m, st, f0);
    }
    compact_faces_kernel<<<1, 256, 0, st>>>(detector_boxes(p->det), detector_counts(p->det), batch, p->max_faces, p->faces_dev, p->n_dev);
    count_launch();
    FRB_CUDA(cudaGetLastError());
    FRB_CUDA(cudaMemcpyAsync(p->h_n, p->n_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaMemcpyAsync(p->h_boxes, detector_boxes(p->det), sizeof(FrBbox) * batch * p->max_faces, cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaMemcpyAsync(p->h_counts, detector_counts(p->det), sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaEventRecord(p->count_ready, st));
    // speculative first chunk: enqueued before the host knows the face count, so the GPU never waits for the host in the middle
    const int slots = batch * p->max_faces;
    const int s0 = std::min({p->emb_batch, slots, p->spec > 0 ? p->spec : slots});
    embed_chunk(p, fdev, static_cast<int>(row_bytes), 0, s0, st);
    FRB_CUDA(cudaEventSynchronize(p->count_ready));  // the GPU is busy with the chunk above while the host learns the count
    const int n = *p->h_n;
    p->spec = std::max(32, (n + 31) / 32 * 32);
    for (int beg = s0; beg < n; beg += p->emb_batch)  // chunked like ArcFaceIR50::forward (src/arcface.cpp:177-185), rows i -> i
        embed_chunk(p, fdev, static_cast<int>(row_bytes), beg, std::min(p->emb_batch, n - beg), st);
    if (n > 0) {
        if (embeddings_host) FRB_CUDA(cudaMemcpyAsync(embeddings_host, p->embeds_dev, sizeof(float) * n * 512, cudaMemcpyDeviceToHost, st));
        if (p->gal && fr_gallery_rows(p->gal) > 0) {
            const int rc = fr_gallery_topk_dev(p->gal, p->embeds_dev, n, 1, p->score_dev, reinterpret_cast<int64_t*>(p->idx_dev), st);
            if (rc != FR_OK) throw CudaError{std::string("pipeline search failed: ") + fr_last_error()};
            FRB_CUDA(cudaMemcpyAsync(p->h_score, p->score_dev, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
            FRB_CUDA(cudaMemcpyAsync(p->h_idx, p->idx_dev, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
        } else {
            std::fill(p->h_idx, p->h_idx + n, -1LL);
            std::fill(p->h_score, p->h_score + n, 0.f);
        }
    }
    FRB_CUDA(cudaStreamSynchronize(st));
    return n;
}

}  // namespace

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
        try {
            FRB_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
            FRB_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
            p->h2d_done.resize((p->max_batch + p->sub_batch - 1) / p->sub_batch);
            for (auto& e : p->h2d_done) FRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            FRB_CUDA(cudaEventCreateWithFlags(&p->count_ready, cudaEventDisableTiming));
            FRB_CUDA(cudaMalloc(&p->faces_dev, sizeof(FaceRef) * slots));
            FRB_CUDA(cudaMalloc(&p->n_dev, sizeof(int)));
            FRB_CUDA(cudaMalloc(&p->embeds_dev, sizeof(float) * slots * 512));
            FRB_CUDA(cudaMalloc(&p->score_dev, sizeof(float) * slots));
            FRB_CUDA(cudaMalloc(&p->idx_dev, sizeof(long long) * slots));
            FRB_CUDA(cudaMallocHost(&p->h_n, sizeof(int)));
            FRB_CUDA(cudaMallocHost(&p->h_boxes, sizeof(FrBbox) * slots));
            FRB_CUDA(cudaMallocHost(&p->h_counts, sizeof(int) * p->max_batch));
            FRB_CUDA(cudaMallocHost(&p->h_score, sizeof(float) * slots));
            FRB_CUDA(cudaMallocHost(&p->h_idx, sizeof(long long) * slots));
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
    cudaFree(p->faces_dev);
    cudaFree(p->n_dev);
    cudaFree(p->embeds_dev);
    cudaFree(p->score_dev);
    cudaFree(p->idx_dev);
    cudaFreeHost(p->h_n);
    cudaFreeHost(p->h_boxes);
    cudaFreeHost(p->h_counts);
    cudaFreeHost(p->h_score);
    cudaFreeHost(p->h_idx);
    for (auto e : p->h2d_done)
        if (e) cudaEventDestroy(e);
    if (p->count_ready) cudaEventDestroy(p->count_ready);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (prev >= 0) cudaSetDevice(prev);
    delete p;
}

int fr_pipeline_run(FrPipeline* p, const uint8_t* frames, int stride, int batch, FrBbox* boxes, int* counts, int64_t* top1_idx,
                    float* top1_score, float* embeddings) {
    return guarded([&] {
        if (!p || !frames || !boxes || !counts) throw ArgError{"null argument"};
        if (batch < 1 || batch > p->max_batch) throw ArgError{"batch out of range (1..max_batch)"};
        if (stride < p->frame_w * 3) throw ArgError{"stride smaller than frame_w * 3"};
        DeviceGuard dg(p->device);
        // per-frame face slots: face j of frame f -> slot f * max_faces + j; embeddings are compact (face order)
        std::vector<float> compact;
        float* emb_tmp = nullptr;
        if (embeddings) {
            compact.resize(static_cast<size_t>(batch) * p->max_faces * 512);
            emb_tmp = compact.data();
        }
        int n = 0;
        try {
            n = run_batch(p, frames, stride, batch, emb_tmp);
        } catch (...) {  // leave no work in flight that still reads the caller's buffers
            cudaStreamSynchronize(p->copy_stream);
            cudaStreamSynchronize(p->stream);
            throw;
        }
        std::copy(p->h_boxes, p->h_boxes + static_cast<size_t>(batch) * p->max_faces, boxes);
        std::copy(p->h_counts, p->h_counts + batch, counts);
        int k = 0;
        for (int f = 0; f < batch; ++f)
            for (int j = 0; j < p->max_faces; ++j) {
                const size_t slot = static_cast<size_t>(f) * p->max_faces + j;
                const bool live = j < counts[f];
                if (top1_idx) top1_idx[slot] = live ? p->h_idx[k] : -1;
                if (top1_score) top1_score[slot] = live ? p->h_score[k] : 0.f;
                if (embeddings) {
                    if (live) std::copy(emb_tmp + static_cast<size_t>(k) * 512, emb_tmp + static_cast<size_t>(k + 1) * 512, embeddings + slot * 512);
                    else std::fill(embeddings + slot * 512, embeddings + (slot + 1) * 512, 0.f);
                }
                if (live) ++k;
            }
        if (k != n) throw CudaError{"pipeline: device face count disagrees with the per-frame counts"};
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
