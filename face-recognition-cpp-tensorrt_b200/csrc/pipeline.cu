// End-to-end pipeline on one GPU: detect -> crop + bicubic resize -> embed -> search, one stream, no host round trip except
// the (tiny) box list. Replaces the body of the reference's /inference handler (/root/reference src/app.cpp:293-352):
// detector.findFace (src/retinaface.cpp:147-152) -> recognizer.forward (src/arcface.cpp:166-187: getCroppedFaces :3-17,
// preprocessFaces :116-129, doInference :139-148) -> featureMatching + getOutputs (:189-217).
#include <algorithm>
#include <memory>
#include <vector>

#include "common.h"

// internal hooks (detector.cu / embedder.cu)
namespace frb {
void detector_forward_dev(FrDetector* d, const uint8_t* frames_dev, int stride, int batch, cudaStream_t st);
uint8_t* detector_frames_buffer(FrDetector* d);
const FrBbox* detector_boxes(const FrDetector* d);
const int* detector_counts(const FrDetector* d);
void detector_dims(const FrDetector* d, int* frame_h, int* frame_w, int* max_batch, int* max_faces, int* device);
uint8_t* embedder_u8_input(FrEmbedder* e);
float* embedder_output(FrEmbedder* e);
int embedder_max_batch(const FrEmbedder* e);
int embedder_device(const FrEmbedder* e);
void embedder_forward_u8(FrEmbedder* e, int batch, cudaStream_t st);
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
                                                          const FaceRef* __restrict__ faces, const int* __restrict__ n_faces_dev, int n_faces,
                                                          uint8_t* __restrict__ out) {
    constexpr int D = 112;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int face = blockIdx.y;
    if (n_faces_dev) n_faces = min(n_faces, *n_faces_dev);  // device-side face list (fr_pipeline_run): the count lives on the device
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

}  // namespace

struct FrPipeline {
    FrDetector* det = nullptr;
    FrEmbedder* emb = nullptr;
    FrGallery* gal = nullptr;
    int device = 0, frame_h = 0, frame_w = 0, max_batch = 0, max_faces = 0, emb_batch = 0;
    cudaStream_t stream = nullptr;
    FaceRef* faces_dev = nullptr;
    float* embeds_dev = nullptr;     // max_batch * max_faces x 512
    float* score_dev = nullptr;
    long long* idx_dev = nullptr;
    std::vector<FrBbox> h_boxes;
    std::vector<int> h_counts;
    std::vector<FaceRef> h_faces;
    std::vector<float> h_score;
    std::vector<long long> h_idx;
};

namespace {

void crop_into_embedder(FrPipeline* p, const uint8_t* frames_dev, int stride, const FaceRef* faces_dev, int n, cudaStream_t st) {
    dim3 grid((112 * 112 + 255) / 256, n);
    crop_resize_kernel<<<grid, 256, 0, st>>>(frames_dev, p->frame_h, p->frame_w, stride, faces_dev, nullptr, n, embedder_u8_input(p->emb));
    count_launch();
    FRB_CUDA(cudaGetLastError());
}

// frames already on the device. Fills the host vectors of the pipeline; returns the number of faces.
int run_device(FrPipeline* p, const uint8_t* frames_dev, int stride, int batch, float* embeddings_host) {
    cudaStream_t st = p->stream;
    detector_forward_dev(p->det, frames_dev, stride, batch, st);
    FRB_CUDA(cudaMemcpyAsync(p->h_boxes.data(), detector_boxes(p->det), sizeof(FrBbox) * batch * p->max_faces, cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaMemcpyAsync(p->h_counts.data(), detector_counts(p->det), sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaStreamSynchronize(st));
    p->h_faces.clear();
    for (int f = 0; f < batch; ++f)
        for (int j = 0; j < p->h_counts[f]; ++j) {
            const FrBbox& b = p->h_boxes[static_cast<size_t>(f) * p->max_faces + j];
            p->h_faces.push_back(FaceRef{f, b.x1, b.y1, b.x2, b.y2});
        }
    const int n = static_cast<int>(p->h_faces.size());
    if (n == 0) return 0;
    FRB_CUDA(cudaMemcpyAsync(p->faces_dev, p->h_faces.data(), sizeof(FaceRef) * n, cudaMemcpyHostToDevice, st));
    for (int beg = 0; beg < n; beg += p->emb_batch) {  // chunked like ArcFaceIR50::forward (src/arcface.cpp:177-185), rows i -> i
        const int m = std::min(p->emb_batch, n - beg);
        crop_into_embedder(p, frames_dev, stride, p->faces_dev + beg, m, st);
        embedder_forward_u8(p->emb, m, st);
        FRB_CUDA(cudaMemcpyAsync(p->embeds_dev + static_cast<size_t>(beg) * 512, embedder_output(p->emb), sizeof(float) * m * 512,
                                 cudaMemcpyDeviceToDevice, st));
    }
    if (embeddings_host) FRB_CUDA(cudaMemcpyAsync(embeddings_host, p->embeds_dev, sizeof(float) * n * 512, cudaMemcpyDeviceToHost, st));
    if (p->gal && fr_gallery_rows(p->gal) > 0) {
        const int rc = fr_gallery_topk_dev(p->gal, p->embeds_dev, n, 1, p->score_dev, reinterpret_cast<int64_t*>(p->idx_dev), st);
        if (rc != FR_OK) throw CudaError{std::string("pipeline search failed: ") + fr_last_error()};
        FRB_CUDA(cudaMemcpyAsync(p->h_score.data(), p->score_dev, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
        FRB_CUDA(cudaMemcpyAsync(p->h_idx.data(), p->idx_dev, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    } else {
        std::fill(p->h_idx.begin(), p->h_idx.begin() + n, -1LL);
        std::fill(p->h_score.begin(), p->h_score.begin() + n, 0.f);
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
        p->emb_batch = embedder_max_batch(emb);
        DeviceGuard dg(p->device);
        const size_t slots = static_cast<size_t>(p->max_batch) * p->max_faces;
        try {
            FRB_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
            FRB_CUDA(cudaMalloc(&p->faces_dev, sizeof(FaceRef) * slots));
            FRB_CUDA(cudaMalloc(&p->embeds_dev, sizeof(float) * slots * 512));
            FRB_CUDA(cudaMalloc(&p->score_dev, sizeof(float) * slots));
            FRB_CUDA(cudaMalloc(&p->idx_dev, sizeof(long long) * slots));
        } catch (...) {
            fr_pipeline_destroy(p.release());
            throw;
        }
        p->h_boxes.resize(slots);
        p->h_counts.resize(p->max_batch);
        p->h_score.resize(slots);
        p->h_idx.resize(slots);
        *out = p.release();
    });
}

void fr_pipeline_destroy(FrPipeline* p) {
    if (!p) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    cudaFree(p->faces_dev);
    cudaFree(p->embeds_dev);
    cudaFree(p->score_dev);
    cudaFree(p->idx_dev);
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
        uint8_t* fdev = detector_frames_buffer(p->det);
        FRB_CUDA(cudaMemcpy2DAsync(fdev, static_cast<size_t>(p->frame_w) * 3, frames, stride, static_cast<size_t>(p->frame_w) * 3,
                                   static_cast<size_t>(batch) * p->frame_h, cudaMemcpyHostToDevice, p->stream));
        // per-frame face slots: face j of frame f -> slot f * max_faces + j; embeddings are compact (face order)
        std::vector<float> compact;
        float* emb_tmp = nullptr;
        if (embeddings) {
            compact.resize(static_cast<size_t>(batch) * p->max_faces * 512);
            emb_tmp = compact.data();
        }
        const int n = run_device(p, fdev, p->frame_w * 3, batch, emb_tmp);
        std::copy(p->h_boxes.begin(), p->h_boxes.begin() + static_cast<size_t>(batch) * p->max_faces, boxes);
        std::copy(p->h_counts.begin(), p->h_counts.begin() + batch, counts);
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
        (void)n;
    });
}

/* getCroppedFaces + preprocessFaces + doInference for caller-provided boxes of ONE frame (ArcFaceIR50::forward,
 * src/arcface.cpp:166-187). crops_u8 (optional): n x 112 x 112 x 3 BGR = CroppedFace::face. Host buffers. */
int fr_embedder_run_boxes(FrEmbedder* e, const uint8_t* frame, int frame_h, int frame_w, int stride, const FrBbox* boxes, int n, float* out512,
                          uint8_t* crops_u8) {
    return guarded([&] {
        if (!e || !frame || !boxes || !out512) throw ArgError{"null argument"};
        if (n < 1) throw ArgError{"no boxes"};
        if (frame_h < 1 || frame_w < 1 || stride < frame_w * 3) throw ArgError{"bad frame geometry"};
        DeviceGuard dg(embedder_device(e));
        cudaStream_t st = nullptr;
        FRB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        uint8_t* fdev = nullptr;
        FaceRef* faces_dev = nullptr;
        try {
            FRB_CUDA(cudaMalloc(&fdev, static_cast<size_t>(frame_h) * frame_w * 3));
            FRB_CUDA(cudaMalloc(&faces_dev, sizeof(FaceRef) * n));
            FRB_CUDA(cudaMemcpy2DAsync(fdev, static_cast<size_t>(frame_w) * 3, frame, stride, static_cast<size_t>(frame_w) * 3, frame_h,
                                       cudaMemcpyHostToDevice, st));
            std::vector<FaceRef> faces(n);
            for (int i = 0; i < n; ++i) faces[i] = FaceRef{0, boxes[i].x1, boxes[i].y1, boxes[i].x2, boxes[i].y2};
            FRB_CUDA(cudaMemcpyAsync(faces_dev, faces.data(), sizeof(FaceRef) * n, cudaMemcpyHostToDevice, st));
            const int mb = embedder_max_batch(e);
            for (int beg = 0; beg < n; beg += mb) {
                const int m = std::min(mb, n - beg);
                dim3 grid((112 * 112 + 255) / 256, m);
                crop_resize_kernel<<<grid, 256, 0, st>>>(fdev, frame_h, frame_w, frame_w * 3, faces_dev + beg, nullptr, m, embedder_u8_input(e));
                count_launch();
                if (crops_u8)
                    FRB_CUDA(cudaMemcpyAsync(crops_u8 + static_cast<size_t>(beg) * 112 * 112 * 3, embedder_u8_input(e),
                                             static_cast<size_t>(m) * 112 * 112 * 3, cudaMemcpyDeviceToHost, st));
                embedder_forward_u8(e, m, st);
                FRB_CUDA(cudaMemcpyAsync(out512 + static_cast<size_t>(beg) * 512, embedder_output(e), sizeof(float) * m * 512, cudaMemcpyDeviceToHost, st));
            }
            FRB_CUDA(cudaStreamSynchronize(st));
        } catch (...) {
            cudaStreamSynchronize(st);
            cudaFree(fdev);
            cudaFree(faces_dev);
            cudaStreamDestroy(st);
            throw;
        }
        cudaFree(fdev);
        cudaFree(faces_dev);
        cudaStreamDestroy(st);
    });
}

}  // extern "C"
