// ORACLE / TEST INFRASTRUCTURE ONLY — never linked into libfr_b200.
// extern "C" handle around the reference's own RetinaFace class, compiled VERBATIM from /root/reference/src/retinaface.cpp
// (+ common.cpp) by oracle/build_ref.py against the stub headers in oracle/stubs/. It pins
//     RetinaFace::RetinaFace               sizes and scales                 src/retinaface.cpp:3-29
//     RetinaFace::create_anchor_retinaface                                  src/retinaface.cpp:210-240
//     RetinaFace::postprocessing           decode/threshold/rescale/clip    src/retinaface.cpp:154-208
//     RetinaFace::nms                                                        src/retinaface.cpp:248-271
// as the reference itself executes them: oracle/retina_post.c (the restatement) and det_decode_nms_kernel are tested against
// this library. The network (TensorRT) and preprocess (OpenCV) halves of the class are not exercisable here.
//
// `#define private public` only opens the class for this accessor translation unit; retinaface.cpp itself is compiled untouched
// (access specifiers do not change layout or code generation).
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define private public
#include "retinaface.h"
#undef private

// The constructor allocates device buffers and a stream through the CUDA runtime (src/retinaface.cpp:84-103). The oracle has no
// GPU work to do, so the few runtime entry points the class touches are satisfied here with host memory; the library is linked
// with -Bsymbolic and without libcudart so these definitions are the ones retinaface.cpp / common.cpp bind to.
extern "C" {
cudaError_t cudaMalloc(void **p, size_t n) {
    *p = std::malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) {
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t *s) {
    *s = nullptr;
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::memcpy(dst, src, n);
    return cudaSuccess;
}
const char *cudaGetErrorString(cudaError_t) { return "stubbed CUDA runtime (oracle build)"; }
}

extern "C" {

// engine_file must exist (loadEngine reads it, src/retinaface.cpp:32-48); its content is ignored by the stub runtime.
void *ref_retina_new(const char *engine_file, int frame_w, int frame_h, int in_c, int in_h, int in_w, int max_batch, int max_faces,
                     float nms_thr, float bbox_thr) {
    try {
        TRTLogger logger;
        std::vector<std::string> outs = {"output_det0", "output_det1"};
        std::vector<int> shape = {in_c, in_h, in_w};
        return new RetinaFace(logger, engine_file, frame_w, frame_h, "input_det", outs, shape, max_batch, max_faces, nms_thr, bbox_thr);
    } catch (...) {
        return nullptr;
    }
}
void ref_retina_free(void *h) {
    try {
        delete static_cast<RetinaFace *>(h);
    } catch (...) {
    }
}
int ref_retina_output_size_base(void *h) { return static_cast<RetinaFace *>(h)->m_OUTPUT_SIZE_BASE; }
void ref_retina_scales(void *h, float *scale_h, float *scale_w) {
    *scale_h = static_cast<RetinaFace *>(h)->m_scale_h;
    *scale_w = static_cast<RetinaFace *>(h)->m_scale_w;
}
// anchors for a w x h network input: out[cap][4] = {cx, cy, sx, sy}; returns the count
int ref_retina_anchors(void *h, int w, int hh, float *out, int cap) {
    std::vector<anchorBox> a;
    static_cast<RetinaFace *>(h)->create_anchor_retinaface(a, w, hh);
    const int n = static_cast<int>(a.size());
    for (int i = 0; i < n && i < cap; ++i) {
        out[4 * i + 0] = a[i].cx;
        out[4 * i + 1] = a[i].cy;
        out[4 * i + 2] = a[i].sx;
        out[4 * i + 3] = a[i].sy;
    }
    return n;
}
// RetinaFace::postprocessing on caller-provided head outputs (bbox [A][4], conf [A][2]); returns m_outputBbox (count, <= cap copied)
int ref_retina_postprocess(void *h, float *bbox, float *conf, Bbox *out, int cap) {
    RetinaFace *r = static_cast<RetinaFace *>(h);
    r->postprocessing(bbox, conf);
    const int n = static_cast<int>(r->m_outputBbox.size());
    for (int i = 0; i < n && i < cap; ++i) out[i] = r->m_outputBbox[i];
    return n;
}
// RetinaFace::nms on caller-provided (already sorted) boxes, in place; returns the survivor count
int ref_retina_nms(void *h, Bbox *boxes, int n, float thr) {
    std::vector<Bbox> v(boxes, boxes + n);
    static_cast<RetinaFace *>(h)->nms(v, thr);
    for (size_t i = 0; i < v.size(); ++i) boxes[i] = v[i];
    return static_cast<int>(v.size());
}
}
