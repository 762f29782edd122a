// RetinaFace mobile0.25 detector: host side of the C ABI ("Detector" section of include/fr_b200.h).
// Replaces RetinaFace (/root/reference src/retinaface.{h,cpp}): preprocess (:106-136), the TensorRT engine (:138-145; network =
// conversion/retina/models/{net,retinaface_trim,retinaface}.py) and postprocessing + nms (:154-271).
//
// Plan per batch: stem (u8 canvas -> 8ch, fused mean subtraction) -> 13 x { depthwise 3x3 ; pointwise 1x1 } -> FPN laterals
// (the top-down nearest-x2 add is fused into the lateral's epilogue) + merges -> 3 x SSH -> one fused 64->32 head GEMM per level
// -> bias + softmax + anchor-major scatter -> decode + NMS, one block per image. Convs with >= 64 input channels are
// conv_gemm_kernel launches (tcgen05); the rest are the CUDA-core kernels of det_kernels.cuh.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "common.h"
#include "conv_kernels.cuh"
#include "det_kernels.cuh"
#include "det_dwpw_gemm.cuh"
#include "weights.h"

using namespace frb;

namespace {

struct DBuf {
    __half* p = nullptr;
    int C = 0;
    Geo g{0, 0};
    CUtensorMap tmap{};
    bool has_map = false;
};

enum StepKind { kDw, kDwPwSmall, kDwPwGemm, kGemm, kC16 };

struct DStep {
    StepKind kind;
    // dw / pw_small / c16
    const __half* in = nullptr;
    __half* out = nullptr;
    Geo gi{0, 0}, go{0, 0};
    int stride = 1, cin = 0, cout = 0, ld_out = 0;
    const float* w = nullptr;
    const float* b = nullptr;
    DwW dww{};                  // fused small blocks: host copy of the depthwise weights / bias (kernel parameter, see StemW)
    const float* w2 = nullptr;  // fused blocks: pointwise weights / bias; c16: second conv on the same input
    const float* b2 = nullptr;
    __half* out2 = nullptr;
    int ld_out2 = 0;
    // gemm
    CUtensorMap ta{}, tb{};
    ConvGemmParams prm{};
    int bn = 64;
    bool heads = false;  // gemm: HEADS epilogue
    int lane = 0;        // 0: main chain; 1, 2: side chains (SSH of level 3 / level 2) that only depend on what the main chain produced so far
    DwPwGemmParams dp{};  // fused depthwise + pointwise GEMM (tb = pointwise weights)
};

}  // namespace

struct FrDetector {
    int device = 0, sms = 0, kind = 0;
    int net_h = 0, net_w = 0, frame_h = 0, frame_w = 0, max_batch = 0, max_faces = 0, anchors = 0;
    float nms_thr = 0.4f, bbox_thr = 0.6f;
    bool landmarks = false;
    cudaStream_t stream = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr};      // side chains of the network (see build_plan)
    cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
    std::vector<void*> allocs;
    float *stem_w = nullptr, *stem_b = nullptr;
    StemW stem_cw{};                 // host copy of the stem weights, passed to det_stem_kernel as a kernel parameter
    uint8_t* frames_dev = nullptr;   // max_batch x frame_h x frame_w x 3
    uint8_t* canvas_dev = nullptr;   // max_batch x net_h x net_w x 3 (only when frame size != network size)
    float* chw_dev = nullptr;        // fr_detector_net input
    __half* a0 = nullptr;            // stem output
    Geo g[6];
    std::vector<DStep> steps;
    float *loc = nullptr, *conf = nullptr, *landm = nullptr;
    DetCand* cand = nullptr;
    int* n_cand = nullptr;           // per image: candidates appended by det_decode_kernel (zero between batches)
    FrBbox* boxes = nullptr;
    int* counts = nullptr;
    float* out_landm = nullptr;
    int* out_ids = nullptr;
    GraphCache graphs;
};

namespace {

template <class T>
T* dalloc(FrDetector* d, size_t count, bool zero) {
    T* p = nullptr;
    FRB_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    d->allocs.push_back(p);
    if (zero) FRB_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), d->stream));
    return p;
}
template <class T>
T* dupload(FrDetector* d, const HostTensor& t) {
    T* p = dalloc<T>(d, static_cast<size_t>(t.numel()), false);
    FRB_CUDA(cudaMemcpyAsync(p, t.data, t.nbytes, cudaMemcpyHostToDevice, d->stream));
    return p;
}

DBuf mkbuf(FrDetector* d, Geo g, int C) {
    DBuf b;
    b.g = g;
    b.C = C;
    const size_t rows = static_cast<size_t>(d->max_batch) * g.HpWp();
    b.p = dalloc<__half>(d, rows * C, true);  // pads are zeroed once and never written
    if (C >= 64) {
        b.tmap = make_tmap_2d_f16(b.p, rows, static_cast<uint64_t>(C), 128, 64);
        b.has_map = true;
    }
    return b;
}

// FR_DET_LANES=0: the SSH blocks of levels 2 and 3 run in the main chain instead of as parallel graph branches (A/B)
const bool g_det_lanes = std::getenv("FR_DET_LANES") == nullptr || std::atoi(std::getenv("FR_DET_LANES")) != 0;
// FR_DET_FUSED_GEMM=0: conv_dw blocks with >= 64 input channels run as dw3x3_kernel + conv_gemm_kernel instead of dwpw_gemm_kernel (A/B)
const bool g_fused_dwpw = std::getenv("FR_DET_FUSED_GEMM") != nullptr && std::atoi(std::getenv("FR_DET_FUSED_GEMM")) != 0;

template <int BN>
void launch_dwpw_gemm(const DStep& s, int batch, cudaStream_t st) {
    DwPwGemmParams p = s.dp;
    p.P = batch * s.go.HpWp();
    launch_k(dwpw_gemm_kernel<BN>, dim3((p.P + kConvBM - 1) / kConvBM), dim3(256), DwPwGemmCfg<BN>::smem_bytes(p.cin), st, true, s.tb, p);
    count_launch();
}

void build_plan(FrDetector* d, const WeightFile& wf) {
    auto f32 = [&](const std::string& n, int64_t numel) { return dupload<float>(d, wf.get(n, 0, numel)); };
    auto f16 = [&](const std::string& n, int64_t numel) { return dupload<__half>(d, wf.get(n, 1, numel)); };
    const int B = d->max_batch;
    for (int k = 1; k <= 5; ++k) d->g[k] = Geo{d->net_h >> k, d->net_w >> k};
    d->stem_w = f32("stem.w", 8 * 27);
    d->stem_b = f32("stem.b", 8);
    {
        const float* hw = reinterpret_cast<const float*>(wf.get("stem.w", 0, 8 * 27).data);
        const float* hb = reinterpret_cast<const float*>(wf.get("stem.b", 0, 8).data);
        for (int i = 0; i < 27 * 8; ++i) d->stem_cw.w[i % 27][i / 27] = hw[i];  // file: [output channel][27]
        for (int n = 0; n < 8; ++n) d->stem_cw.b[n] = hb[n];
    }
    d->frames_dev = dalloc<uint8_t>(d, static_cast<size_t>(B) * d->frame_h * d->frame_w * 3, false);
    d->chw_dev = dalloc<float>(d, static_cast<size_t>(B) * 3 * d->net_h * d->net_w, false);

    DBuf cur = mkbuf(d, d->g[1], 8);
    d->a0 = cur.p;

    auto add_gemm = [&](const DBuf& in, int cin, int taps, int cout, const std::string& wname, const std::string& bname, bool relu,
                        __half* out, int ld_out, const __half* res, int res_mode, float* out_f32) {
        DStep s{};
        s.kind = kGemm;
        s.bn = cout >= 128 ? 128 : cout;
        __half* w = f16(wname, static_cast<int64_t>(cout) * taps * cin);
        s.ta = in.tmap;
        s.tb = make_tmap_2d_f16(w, cout, static_cast<uint64_t>(taps) * cin, s.bn, 64);
        s.go = in.g;
        s.prm.H = in.g.H;
        s.prm.W = in.g.W;
        s.prm.cin_blocks = cin / 64;
        s.prm.taps = taps;
        s.prm.cout = cout;
        s.prm.kb_per_split = taps * (cin / 64);
        if (out_f32) {
            s.prm.partial = out_f32;
        } else {
            s.prm.bias = f32(bname, cout);
            s.prm.relu = relu ? 1 : 0;
            s.prm.out = out;
            s.prm.ld_out = ld_out;
            s.prm.res = res;
            s.prm.res_mode = res_mode;
            s.prm.ld_res = 64;
        }
        d->steps.push_back(s);
    };

    // ---- MobileNetV1-0.25 body (net.py:102-124): 13 conv_dw blocks
    const int dw_cin[13] = {8, 16, 32, 32, 64, 64, 128, 128, 128, 128, 128, 128, 256};
    const int dw_cout[13] = {16, 32, 32, 64, 64, 128, 128, 128, 128, 128, 128, 256, 256};
    const int dw_stride[13] = {1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1};
    int level = 1;
    DBuf c1{}, c2{}, c3{};
    for (int n = 0; n < 13; ++n) {
        const std::string id = std::to_string(n + 1);
        if (dw_stride[n] == 2) ++level;
        DBuf pwo = mkbuf(d, d->g[level], dw_cout[n]);
        if (dw_cin[n] < 64) {
            // fused depthwise + pointwise on CUDA cores (dwpw_small_kernel)
            DStep s{};
            s.kind = kDwPwSmall;
            s.in = cur.p;
            s.out = pwo.p;
            s.gi = cur.g;
            s.go = pwo.g;
            s.stride = dw_stride[n];
            s.cin = dw_cin[n];
            s.cout = dw_cout[n];
            s.w = f32("dw" + id + ".w", 9 * dw_cin[n]);
            s.b = f32("dw" + id + ".b", dw_cin[n]);
            {
                const float* hw = reinterpret_cast<const float*>(wf.get("dw" + id + ".w", 0, 9 * dw_cin[n]).data);  // [tap][channel]
                const float* hb = reinterpret_cast<const float*>(wf.get("dw" + id + ".b", 0, dw_cin[n]).data);
                for (int t = 0; t < 9; ++t)
                    for (int c = 0; c < dw_cin[n]; ++c) s.dww.w[t][c] = hw[t * dw_cin[n] + c];
                for (int c = 0; c < dw_cin[n]; ++c) s.dww.b[c] = hb[c];
            }
            s.w2 = f32("pw" + id + ".w", static_cast<int64_t>(dw_cin[n]) * dw_cout[n]);
            s.b2 = f32("pw" + id + ".b", dw_cout[n]);
            d->steps.push_back(s);
        } else if (g_fused_dwpw) {
            // fused depthwise (CUDA cores, straight into the tensor core's shared-memory operand) + pointwise GEMM (dwpw_gemm_kernel)
            DStep s{};
            s.kind = kDwPwGemm;
            s.bn = dw_cout[n];
            __half* w = f16("pw" + id + ".w", static_cast<int64_t>(dw_cout[n]) * dw_cin[n]);
            s.tb = make_tmap_2d_f16(w, dw_cout[n], dw_cin[n], dw_cout[n], 64);
            s.go = pwo.g;
            s.dp.in = cur.p;
            s.dp.gi = cur.g;
            s.dp.go = pwo.g;
            s.dp.stride = dw_stride[n];
            s.dp.cin = dw_cin[n];
            s.dp.dw_w = f32("dw" + id + ".w", 9 * dw_cin[n]);
            s.dp.dw_b = f32("dw" + id + ".b", dw_cin[n]);
            s.dp.pw_b = f32("pw" + id + ".b", dw_cout[n]);
            s.dp.out = pwo.p;
            d->steps.push_back(s);
        } else {
            DBuf dwo = mkbuf(d, d->g[level], dw_cin[n]);
            DStep s{};
            s.kind = kDw;
            s.in = cur.p;
            s.out = dwo.p;
            s.gi = cur.g;
            s.go = dwo.g;
            s.stride = dw_stride[n];
            s.cin = dw_cin[n];
            s.w = f32("dw" + id + ".w", 9 * dw_cin[n]);
            s.b = f32("dw" + id + ".b", dw_cin[n]);
            d->steps.push_back(s);
            add_gemm(dwo, dw_cin[n], 1, dw_cout[n], "pw" + id + ".w", "pw" + id + ".b", true, pwo.p, 0, nullptr, kResNone, nullptr);
        }
        cur = pwo;
        if (n == 4) c1 = cur;
        if (n == 10) c2 = cur;
        if (n == 12) c3 = cur;
    }
    // ---- FPN (net.py:68-98): laterals, top-down nearest x2 add fused into the lateral epilogue, 3x3 merges
    DBuf o3 = mkbuf(d, d->g[5], 64), m2in = mkbuf(d, d->g[4], 64), o2 = mkbuf(d, d->g[4], 64), m1in = mkbuf(d, d->g[3], 64),
         o1 = mkbuf(d, d->g[3], 64);
    add_gemm(c3, 256, 1, 64, "fpn.output3.w", "fpn.output3.b", true, o3.p, 0, nullptr, kResNone, nullptr);
    add_gemm(c2, 128, 1, 64, "fpn.output2.w", "fpn.output2.b", true, m2in.p, 0, o3.p, kResUpsample, nullptr);
    add_gemm(m2in, 64, 9, 64, "fpn.merge2.w", "fpn.merge2.b", true, o2.p, 0, nullptr, kResNone, nullptr);
    add_gemm(c1, 64, 1, 64, "fpn.output1.w", "fpn.output1.b", true, m1in.p, 0, o2.p, kResUpsample, nullptr);
    add_gemm(m1in, 64, 9, 64, "fpn.merge1.w", "fpn.merge1.b", true, o1.p, 0, nullptr, kResNone, nullptr);
    // ---- SSH x3 (net.py:40-66) + heads (retinaface_trim.py:89-99 / retinaface.py:37-46)
    d->anchors = (d->g[3].H * d->g[3].W + d->g[4].H * d->g[4].W + d->g[5].H * d->g[5].W) * 2;
    d->loc = dalloc<float>(d, static_cast<size_t>(B) * d->anchors * 4, false);
    d->conf = dalloc<float>(d, static_cast<size_t>(B) * d->anchors * 2, false);
    d->landm = d->landmarks ? dalloc<float>(d, static_cast<size_t>(B) * d->anchors * 10, false) : nullptr;
    const DBuf* fpn[3] = {&o1, &o2, &o3};
    int level_offset = 0;
    std::vector<DStep> ssh_steps[4];  // per level; spliced into the plan below
    for (int lvl = 1; lvl <= 3; ++lvl) {
        const size_t first_step = d->steps.size();
        const DBuf& x = *fpn[lvl - 1];
        const std::string p = "ssh" + std::to_string(lvl) + ".";
        DBuf F = mkbuf(d, x.g, 64), T = mkbuf(d, x.g, 16), U = mkbuf(d, x.g, 16);
        {   // conv3X3 (64 -> 32, into F[0:32]) and conv5X5_1 (64 -> 16, into T) read the same map: ONE 64 -> 48 GEMM with two destinations
            DStep s{};
            s.kind = kGemm;
            s.bn = 48;
            const HostTensor wa = wf.get(p + "a.w", 1, 32ll * 9 * 64), wt = wf.get(p + "t.w", 1, 16ll * 9 * 64);
            const HostTensor ba = wf.get(p + "a.b", 0, 32), bt = wf.get(p + "t.b", 0, 16);
            __half* w = dalloc<__half>(d, 48ull * 9 * 64, false);
            float* b = dalloc<float>(d, 48, false);
            FRB_CUDA(cudaMemcpyAsync(w, wa.data, wa.nbytes, cudaMemcpyHostToDevice, d->stream));
            FRB_CUDA(cudaMemcpyAsync(w + 32ull * 9 * 64, wt.data, wt.nbytes, cudaMemcpyHostToDevice, d->stream));
            FRB_CUDA(cudaMemcpyAsync(b, ba.data, ba.nbytes, cudaMemcpyHostToDevice, d->stream));
            FRB_CUDA(cudaMemcpyAsync(b + 32, bt.data, bt.nbytes, cudaMemcpyHostToDevice, d->stream));
            s.ta = x.tmap;
            s.tb = make_tmap_2d_f16(w, 48, 9ull * 64, 48, 64);
            s.go = x.g;
            s.prm.H = x.g.H;
            s.prm.W = x.g.W;
            s.prm.cin_blocks = 1;
            s.prm.taps = 9;
            s.prm.cout = 48;
            s.prm.kb_per_split = 9;
            s.prm.bias = b;
            s.prm.relu = 1;
            s.prm.out = F.p;
            s.prm.ld_out = 64;
            s.prm.out2 = T.p;
            s.prm.out2_from = 32;
            s.prm.ld_out2 = 16;
            d->steps.push_back(s);
        }
        {   // conv5X5_2 (T -> F[32:48]) and conv7X7_2 (T -> U) in one pass over T; then conv7x7_3 (U -> F[48:64])
            DStep s{};
            s.kind = kC16;
            s.in = T.p;
            s.go = x.g;
            s.w = f32(p + "b.w", 9 * 16 * 16);
            s.b = f32(p + "b.b", 16);
            s.out = F.p + 32;
            s.ld_out = 64;
            s.w2 = f32(p + "u.w", 9 * 16 * 16);
            s.b2 = f32(p + "u.b", 16);
            s.out2 = U.p;
            s.ld_out2 = 16;
            d->steps.push_back(s);
            DStep c{};
            c.kind = kC16;
            c.in = U.p;
            c.go = x.g;
            c.w = f32(p + "c.w", 9 * 16 * 16);
            c.b = f32(p + "c.b", 16);
            c.out = F.p + 48;
            c.ld_out = 64;
            d->steps.push_back(c);
        }
        {   // the three 1x1 heads as one 64 -> 32 GEMM whose epilogue adds the bias, applies the softmax and scatters anchor-major
            DStep s{};
            s.kind = kGemm;
            s.bn = 32;
            s.heads = true;
            __half* w = f16("head" + std::to_string(lvl) + ".w", 32ll * 64);
            s.ta = F.tmap;
            s.tb = make_tmap_2d_f16(w, 32, 64, 32, 64);
            s.go = x.g;
            s.prm.H = x.g.H;
            s.prm.W = x.g.W;
            s.prm.cin_blocks = 1;
            s.prm.taps = 1;
            s.prm.cout = 32;
            s.prm.kb_per_split = 1;
            s.prm.bias = f32("head" + std::to_string(lvl) + ".b", 32);
            s.prm.head_loc = d->loc;
            s.prm.head_conf = d->conf;
            s.prm.head_landm = d->landm;
            s.prm.anchors_total = d->anchors;
            s.prm.level_offset = level_offset;
            d->steps.push_back(s);
        }
        level_offset += x.g.H * x.g.W * 2;
        ssh_steps[lvl].assign(d->steps.begin() + first_step, d->steps.end());
        d->steps.resize(first_step);
    }
    // Order of execution: the SSH + heads of a level only need that level's FPN output, so levels 3 and 2 run as side chains (own
    // streams = parallel branches of the captured graph) next to the rest of the FPN and level 1; everything joins before decode.
    {
        std::vector<DStep> plan;
        for (const DStep& s : d->steps) {
            plan.push_back(s);
            if (s.kind == kGemm && s.prm.out == o3.p)
                for (DStep t : ssh_steps[3]) { t.lane = g_det_lanes ? 1 : 0; plan.push_back(t); }
            if (s.kind == kGemm && s.prm.out == o2.p)
                for (DStep t : ssh_steps[2]) { t.lane = g_det_lanes ? 2 : 0; plan.push_back(t); }
        }
        for (const DStep& t : ssh_steps[1]) plan.push_back(t);
        d->steps.swap(plan);
    }
    d->cand = dalloc<DetCand>(d, static_cast<size_t>(B) * d->anchors, false);
    d->n_cand = dalloc<int>(d, B, true);
    d->boxes = dalloc<FrBbox>(d, static_cast<size_t>(B) * d->max_faces, true);
    d->counts = dalloc<int>(d, B, true);
    d->out_landm = dalloc<float>(d, static_cast<size_t>(B) * d->max_faces * 10, true);
    d->out_ids = dalloc<int>(d, static_cast<size_t>(B) * d->max_faces, true);
    FRB_CUDA(cudaStreamSynchronize(d->stream));
}

template <int BN, bool HEADS = false>
void launch_det_gemm(const DStep& s, int batch, cudaStream_t st) {
    ConvGemmParams prm = s.prm;
    prm.P = batch * s.go.HpWp();
    dim3 grid((prm.P + kConvBM - 1) / kConvBM, prm.cout / BN, 1);
    launch_k(conv_gemm_kernel<BN, false, HEADS>, grid, dim3(kConvThreads), ConvCfg<BN>::kSmemBytes, st, true, s.ta, s.tb, prm);
    count_launch();
}

inline int blocks_for(long long threads, int per_block) { return static_cast<int>((threads + per_block - 1) / per_block); }

// network: canvas (u8, net size) or preprocessed f32 planar input -> loc / conf / landm on the device
void run_net(FrDetector* d, const uint8_t* canvas_dev, int stride_bytes, const float* chw_dev, int batch, cudaStream_t st) {
    NvtxRange nvtx("fr.detect.network");
    const Geo g1 = d->g[1];
    const long long px = static_cast<long long>(batch) * g1.H * g1.W;
    if (canvas_dev)
        launch_k(det_stem_kernel, dim3(blocks_for(px, 256)), dim3(256), 0, st, true, canvas_dev, stride_bytes, batch, d->net_h, d->net_w, d->stem_w,
                 d->stem_b, d->a0, d->stem_cw);
    else
        launch_k(det_stem_f32_kernel, dim3(blocks_for(px, 256)), dim3(256), 0, st, true, chw_dev, batch, d->net_h, d->net_w, d->stem_w, d->stem_b, d->a0);
    count_launch();
    bool forked[2] = {false, false};
    cudaStream_t main_st = st;
    for (const DStep& s : d->steps) {
        st = main_st;
        if (s.lane > 0) {
            const int l = s.lane - 1;
            if (!forked[l]) {  // the side chain starts after everything the main chain has enqueued so far
                FRB_CUDA(cudaEventRecord(d->ev_fork[l], main_st));
                FRB_CUDA(cudaStreamWaitEvent(d->side[l], d->ev_fork[l], 0));
                forked[l] = true;
            }
            st = d->side[l];
        }
        switch (s.kind) {
            case kDw: {
                const long long t = static_cast<long long>(batch) * s.go.H * s.go.W * (s.cin / 8);
                static int occ = 0;  // resident CTAs per SM: the grid is one full wave (grid-stride inside the kernel)
                if (!occ) FRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dw3x3_kernel, 256, 0));
                const int nb = static_cast<int>(std::min<long long>((t + 255) / 256, static_cast<long long>(std::max(occ, 1)) * d->sms));
                launch_k(dw3x3_kernel, dim3(nb), dim3(256), 0, st, true, s.in, s.gi, s.out, s.go, s.stride, s.cin, batch, s.w, s.b);
                count_launch();
                break;
            }
            case kDwPwSmall: {
                const long long t = static_cast<long long>(batch) * s.go.H * s.go.W;
                auto launch = [&](auto kern) {
                    int occ = 1;  // resident CTAs per SM of this instantiation: the grid is exactly one full wave (grid-stride inside)
                    FRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
                    const int nb = static_cast<int>(std::min<long long>((t + 255) / 256, static_cast<long long>(std::max(occ, 1)) * d->sms));
                    launch_k(kern, dim3(nb), dim3(256), 0, st, true, s.in, s.gi, s.out, s.go, batch, s.w, s.b, s.w2, s.b2, s.dww);
                };
                if (s.cin == 8 && s.cout == 16 && s.stride == 1) launch(dwpw_small_kernel<8, 16, 1>);
                else if (s.cin == 16 && s.cout == 32 && s.stride == 2) launch(dwpw_small_kernel<16, 32, 2>);
                else if (s.cin == 32 && s.cout == 32 && s.stride == 1) launch(dwpw_small_kernel<32, 32, 1>);
                else if (s.cin == 32 && s.cout == 64 && s.stride == 2) launch(dwpw_small_kernel<32, 64, 2>);
                else throw StateError{"unexpected fused conv_dw shape"};
                count_launch();
                break;
            }
            case kDwPwGemm:
                if (s.bn == 64) launch_dwpw_gemm<64>(s, batch, st);
                else if (s.bn == 128) launch_dwpw_gemm<128>(s, batch, st);
                else if (s.bn == 256) launch_dwpw_gemm<256>(s, batch, st);
                else throw StateError{"unexpected fused conv_dw width"};
                break;
            case kGemm:
                if (s.heads) launch_det_gemm<32, true>(s, batch, st);
                else if (s.bn == 48) launch_det_gemm<48>(s, batch, st);
                else if (s.bn == 64) launch_det_gemm<64>(s, batch, st);
                else if (s.bn == 128) launch_det_gemm<128>(s, batch, st);
                else throw StateError{"unexpected GEMM tile width"};
                break;
            case kC16: {
                const long long t = static_cast<long long>(batch) * s.go.H * s.go.W;
                const int nb = static_cast<int>(std::min<long long>((t + 255) / 256, 2LL * d->sms));  // grid-stride inside the kernel
                if (s.w2)
                    launch_k(conv3x3_c16_kernel<2>, dim3(static_cast<unsigned>(std::min<long long>((t + 127) / 128, 4LL * d->sms))), dim3(128), 0, st, true,
                             s.in, s.go, batch, s.w, s.b, s.out, s.ld_out, s.w2, s.b2, s.out2, s.ld_out2);
                else
                    launch_k(conv3x3_c16_kernel<1>, dim3(nb), dim3(256), 0, st, true, s.in, s.go, batch, s.w, s.b, s.out, s.ld_out,
                             static_cast<const float*>(nullptr), static_cast<const float*>(nullptr), static_cast<__half*>(nullptr), 0);
                count_launch();
                break;
            }
        }
    }
    for (int l = 0; l < 2; ++l)
        if (forked[l]) {
            FRB_CUDA(cudaEventRecord(d->ev_join[l], d->side[l]));
            FRB_CUDA(cudaStreamWaitEvent(main_st, d->ev_join[l], 0));
        }
    FRB_CUDA(cudaGetLastError());
}

// slot0: results of image i go to result slot slot0 + i (the pipeline detects a large batch in sub-batches while later frames are
// still in flight over PCIe)
void run_post(FrDetector* d, const float* loc, const float* conf, const float* landm, int batch, cudaStream_t st, int slot0 = 0) {
    NvtxRange nvtx("fr.detect.decode_nms");
    DetPostParams p;
    p.net_w = d->net_w;
    p.net_h = d->net_h;
    p.frame_w = d->frame_w;
    p.frame_h = d->frame_h;
    p.nms_thr = d->nms_thr;
    p.bbox_thr = d->bbox_thr;
    p.max_faces = d->max_faces;
    p.anchors = d->anchors;
    launch_k(det_decode_kernel, dim3((d->anchors + 255) / 256, batch), dim3(256), 0, st, true, loc, conf, p, d->cand, d->n_cand);
    launch_k(det_nms_kernel, dim3(batch), dim3(256), 0, st, true, landm, p, d->cand, d->n_cand, d->boxes + static_cast<size_t>(slot0) * d->max_faces,
             d->counts + slot0, d->out_landm + static_cast<size_t>(slot0) * d->max_faces * 10, d->out_ids + static_cast<size_t>(slot0) * d->max_faces);
    count_launch(2);
    FRB_CUDA(cudaGetLastError());
}

// network (+ decode/NMS) replayed from a CUDA graph per (canvas pointer, stride, batch, with_post)
void forward_graph(FrDetector* d, const uint8_t* canvas, int cs, int batch, bool with_post, cudaStream_t st, int slot0 = 0) {
    d->graphs.run({reinterpret_cast<uint64_t>(canvas), (static_cast<uint64_t>(cs) << 32) | static_cast<uint64_t>(batch),
                   (with_post ? 1ull : 0ull) | (static_cast<uint64_t>(slot0) << 1)}, st, [&] {
        run_net(d, canvas, cs, nullptr, batch, st);
        if (with_post) run_post(d, d->loc, d->conf, d->landm, batch, st, slot0);
    });
}

void check_det(const FrDetector* d, int batch) {
    if (!d) throw ArgError{"null detector"};
    if (batch < 1 || batch > d->max_batch) throw ArgError{"batch out of range (1..max_batch)"};
}

// frames (host or device) -> letterboxed canvas of the network's size on the device (RetinaFace::preprocess, :106-126);
// returns the canvas pointer and its row stride. When the frame already has the network's size the frame IS the canvas.
const uint8_t* stage_frames(FrDetector* d, const uint8_t* frames, int stride, int batch, bool on_device, int* canvas_stride, cudaStream_t st) {
    if (stride < d->frame_w * 3) throw ArgError{"stride smaller than frame_w * 3"};
    const uint8_t* src = frames;
    int src_stride = stride;
    if (!on_device) {
        FRB_CUDA(cudaMemcpy2DAsync(d->frames_dev, static_cast<size_t>(d->frame_w) * 3, frames, stride, static_cast<size_t>(d->frame_w) * 3,
                                   static_cast<size_t>(batch) * d->frame_h, cudaMemcpyHostToDevice, st));
        src = d->frames_dev;
        src_stride = d->frame_w * 3;
    }
    if (d->frame_h == d->net_h && d->frame_w == d->net_w) {
        *canvas_stride = src_stride;
        return src;
    }
    // letterbox geometry exactly as :111-122 (float scales, truncation to int)
    const float scale_h = static_cast<float>(d->net_h) / d->frame_h, scale_w = static_cast<float>(d->net_w) / d->frame_w;
    int w, h, x, y;
    if (scale_h > scale_w) {
        w = d->net_w;
        h = static_cast<int>(scale_w * d->frame_h);
        x = 0;
        y = (d->net_h - h) / 2;
    } else {
        w = static_cast<int>(scale_h * d->frame_w);
        h = d->net_h;
        x = (d->net_w - w) / 2;
        y = 0;
    }
    if (!d->canvas_dev) d->canvas_dev = dalloc<uint8_t>(d, static_cast<size_t>(d->max_batch) * d->net_h * d->net_w * 3, false);
    const long long px = static_cast<long long>(batch) * d->net_h * d->net_w;
    det_letterbox_kernel<<<blocks_for(px, 256), 256, 0, st>>>(src, d->frame_h, d->frame_w, src_stride, batch, d->net_h, d->net_w, w, h, x, y,
                                                             d->canvas_dev);
    count_launch();
    FRB_CUDA(cudaGetLastError());
    *canvas_stride = d->net_w * 3;
    return d->canvas_dev;
}

void copy_results(FrDetector* d, int batch, FrBbox* boxes, int* counts, float* landmarks, cudaStream_t st) {
    FRB_CUDA(cudaMemcpyAsync(boxes, d->boxes, sizeof(FrBbox) * batch * d->max_faces, cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaMemcpyAsync(counts, d->counts, sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
    if (landmarks) FRB_CUDA(cudaMemcpyAsync(landmarks, d->out_landm, sizeof(float) * batch * d->max_faces * 10, cudaMemcpyDeviceToHost, st));
    FRB_CUDA(cudaStreamSynchronize(st));
}

void copy_raw(FrDetector* d, int batch, float* loc, float* conf, float* landm, cudaStream_t st) {
    const size_t a = static_cast<size_t>(batch) * d->anchors;
    if (loc) FRB_CUDA(cudaMemcpyAsync(loc, d->loc, sizeof(float) * a * 4, cudaMemcpyDeviceToHost, st));
    if (conf) FRB_CUDA(cudaMemcpyAsync(conf, d->conf, sizeof(float) * a * 2, cudaMemcpyDeviceToHost, st));
    if (landm) {
        if (!d->landm) throw StateError{"detector was created without the landmark head"};
        FRB_CUDA(cudaMemcpyAsync(landm, d->landm, sizeof(float) * a * 10, cudaMemcpyDeviceToHost, st));
    }
    FRB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace

// internal hooks for the end-to-end pipeline (csrc/pipeline.cu)
namespace frb {
void detector_forward_dev(FrDetector* d, const uint8_t* frames_dev, int stride, int batch, cudaStream_t st, int slot0) {
    if (slot0 < 0 || slot0 + batch > d->max_batch) throw ArgError{"detector sub-batch outside max_batch"};
    int cs = 0;
    const uint8_t* canvas = stage_frames(d, frames_dev, stride, batch, true, &cs, st);
    forward_graph(d, canvas, cs, batch, true, st, slot0);
}
uint8_t* detector_frames_buffer(FrDetector* d) { return d->frames_dev; }
const FrBbox* detector_boxes(const FrDetector* d) { return d->boxes; }
const int* detector_counts(const FrDetector* d) { return d->counts; }
void detector_dims(const FrDetector* d, int* frame_h, int* frame_w, int* max_batch, int* max_faces, int* device) {
    *frame_h = d->frame_h;
    *frame_w = d->frame_w;
    *max_batch = d->max_batch;
    *max_faces = d->max_faces;
    *device = d->device;
}
}  // namespace frb

namespace frb {
void launch_stretch_resize_u8(const uint8_t* src, int h, int w, int stride_bytes, int out_h, int out_w, uint8_t* dst, cudaStream_t st) {
    const long long px = static_cast<long long>(out_w) * out_h;
    det_letterbox_kernel<<<static_cast<unsigned>((px + 255) / 256), 256, 0, st>>>(src, h, w, stride_bytes, 1, out_h, out_w, out_w, out_h, 0, 0, dst);
    count_launch();
    FRB_CUDA(cudaGetLastError());
}
}  // namespace frb

extern "C" {

int fr_detector_create(const char* weights_path, int net_h, int net_w, int frame_h, int frame_w, int max_batch, int max_faces, float nms_thr,
                       float bbox_thr, int with_landmarks, int device, FrDetector** out) {
    return guarded([&] {
        if (!out) throw ArgError{"out is null"};
        if (net_h < 32 || net_w < 32 || (net_h % 32) || (net_w % 32)) throw ArgError{"network input height/width must be positive multiples of 32"};
        if (frame_h < 1 || frame_w < 1) throw ArgError{"bad frame size"};
        if (max_batch < 1 || max_batch > 1024) throw ArgError{"max_batch out of range (1..1024)"};
        if (max_faces < 1 || max_faces > 1024) throw ArgError{"max_faces out of range (1..1024)"};
        // the detector kernels index pixels (x channel chunks) of a batch with 32-bit arithmetic
        if (static_cast<long long>(max_batch) * (net_h / 2) * (net_w / 2) * 32 >= (1LL << 32)) throw ArgError{"max_batch x network size too large"};
        WeightFile wf = load_weight_file(weights_path);
        if (wf.kind != kKindRetinaTrim && wf.kind != kKindRetinaFull) throw FileError{FR_EFORMAT, "weight file is not a RetinaFace checkpoint"};
        if (with_landmarks && wf.kind != kKindRetinaFull) throw FileError{FR_EFORMAT, "landmarks requested but the checkpoint has no landmark head"};
        std::unique_ptr<FrDetector> d(new FrDetector());
        d->sms = use_device(device);
        d->device = device;
        d->kind = wf.kind;
        d->net_h = net_h;
        d->net_w = net_w;
        d->frame_h = frame_h;
        d->frame_w = frame_w;
        d->max_batch = max_batch;
        d->max_faces = max_faces;
        d->nms_thr = nms_thr;
        d->bbox_thr = bbox_thr;
        d->landmarks = with_landmarks != 0;
        try {
            FRB_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
            for (int l = 0; l < 2; ++l) {
                FRB_CUDA(cudaStreamCreateWithFlags(&d->side[l], cudaStreamNonBlocking));
                FRB_CUDA(cudaEventCreateWithFlags(&d->ev_fork[l], cudaEventDisableTiming));
                FRB_CUDA(cudaEventCreateWithFlags(&d->ev_join[l], cudaEventDisableTiming));
            }
            FRB_CUDA(cudaFuncSetAttribute(dwpw_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwPwGemmCfg<64>::smem_bytes(256)));
            FRB_CUDA(cudaFuncSetAttribute(dwpw_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwPwGemmCfg<128>::smem_bytes(256)));
            FRB_CUDA(cudaFuncSetAttribute(dwpw_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwPwGemmCfg<256>::smem_bytes(256)));
            FRB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<48>::kSmemBytes));
            FRB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<32, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<32>::kSmemBytes));
            FRB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64>::kSmemBytes));
            FRB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<128>::kSmemBytes));
            build_plan(d.get(), wf);
        } catch (...) {
            fr_detector_destroy(d.release());
            throw;
        }
        *out = d.release();
    });
}

void fr_detector_destroy(FrDetector* d) {
    if (!d) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    for (void* p : d->allocs) cudaFree(p);
    for (int l = 0; l < 2; ++l) {
        if (d->side[l]) cudaStreamSynchronize(d->side[l]);
        if (d->ev_fork[l]) cudaEventDestroy(d->ev_fork[l]);
        if (d->ev_join[l]) cudaEventDestroy(d->ev_join[l]);
        if (d->side[l]) cudaStreamDestroy(d->side[l]);
    }
    if (d->stream) cudaStreamDestroy(d->stream);
    if (prev >= 0) cudaSetDevice(prev);
    delete d;
}

int fr_detector_num_anchors(const FrDetector* d) { return d ? d->anchors : FR_EINVAL; }

int fr_detector_run(FrDetector* d, const uint8_t* frames, int stride, int batch, FrBbox* boxes, int* counts, float* landmarks) {
    return guarded([&] {
        check_det(d, batch);
        if (!frames || !boxes || !counts) throw ArgError{"null buffer"};
        DeviceGuard dg(d->device);
        int cs = 0;
        const uint8_t* canvas = stage_frames(d, frames, stride, batch, false, &cs, d->stream);
        forward_graph(d, canvas, cs, batch, true, d->stream);
        copy_results(d, batch, boxes, counts, landmarks, d->stream);
    });
}

int fr_detector_raw(FrDetector* d, const uint8_t* frames, int stride, int batch, float* loc, float* conf, float* landm) {
    return guarded([&] {
        check_det(d, batch);
        if (!frames) throw ArgError{"null buffer"};
        DeviceGuard dg(d->device);
        int cs = 0;
        const uint8_t* canvas = stage_frames(d, frames, stride, batch, false, &cs, d->stream);
        forward_graph(d, canvas, cs, batch, false, d->stream);
        copy_raw(d, batch, loc, conf, landm, d->stream);
    });
}

int fr_detector_net(FrDetector* d, const float* chw, int batch, float* loc, float* conf, float* landm) {
    return guarded([&] {
        check_det(d, batch);
        if (!chw) throw ArgError{"null buffer"};
        DeviceGuard dg(d->device);
        FRB_CUDA(cudaMemcpyAsync(d->chw_dev, chw, sizeof(float) * batch * 3 * d->net_h * d->net_w, cudaMemcpyHostToDevice, d->stream));
        run_net(d, nullptr, 0, d->chw_dev, batch, d->stream);
        copy_raw(d, batch, loc, conf, landm, d->stream);
    });
}

int fr_detector_post(FrDetector* d, const float* loc, const float* conf, const float* landm, int batch, FrBbox* boxes, int* counts,
                     float* landmarks) {
    return guarded([&] {
        check_det(d, batch);
        if (!loc || !conf || !boxes || !counts) throw ArgError{"null buffer"};
        DeviceGuard dg(d->device);
        const size_t a = static_cast<size_t>(batch) * d->anchors;
        FRB_CUDA(cudaMemcpyAsync(d->loc, loc, sizeof(float) * a * 4, cudaMemcpyHostToDevice, d->stream));
        FRB_CUDA(cudaMemcpyAsync(d->conf, conf, sizeof(float) * a * 2, cudaMemcpyHostToDevice, d->stream));
        float* lm_dev = nullptr;
        if (landm) {
            if (!d->landm) {
                d->landm = dalloc<float>(d, static_cast<size_t>(d->max_batch) * d->anchors * 10, false);
            }
            lm_dev = d->landm;
            FRB_CUDA(cudaMemcpyAsync(lm_dev, landm, sizeof(float) * a * 10, cudaMemcpyHostToDevice, d->stream));
        }
        run_post(d, d->loc, d->conf, lm_dev, batch, d->stream);
        copy_results(d, batch, boxes, counts, landmarks, d->stream);
    });
}

int fr_detector_run_dev(FrDetector* d, const uint8_t* frames_dev, int stride, int batch, FrBbox* boxes_dev, int* counts_dev,
                        float* landmarks_dev, void* stream) {
    return guarded([&] {
        check_det(d, batch);
        if (!frames_dev || !boxes_dev || !counts_dev) throw ArgError{"null buffer"};
        DeviceGuard dg(d->device);
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : d->stream;
        int cs = 0;
        const uint8_t* canvas = stage_frames(d, frames_dev, stride, batch, true, &cs, st);
        forward_graph(d, canvas, cs, batch, true, st);
        FRB_CUDA(cudaMemcpyAsync(boxes_dev, d->boxes, sizeof(FrBbox) * batch * d->max_faces, cudaMemcpyDeviceToDevice, st));
        FRB_CUDA(cudaMemcpyAsync(counts_dev, d->counts, sizeof(int) * batch, cudaMemcpyDeviceToDevice, st));
        if (landmarks_dev)
            FRB_CUDA(cudaMemcpyAsync(landmarks_dev, d->out_landm, sizeof(float) * batch * d->max_faces * 10, cudaMemcpyDeviceToDevice, st));
    });
}

}  // extern "C"
