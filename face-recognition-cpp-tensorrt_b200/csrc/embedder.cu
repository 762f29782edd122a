// ArcFace IR-50 / IR-SE-50 embedder: host side of the C ABI ("Embedder" section of include/fr_b200.h).
// Replaces the TensorRT half of ArcFaceIR50 (/root/reference src/arcface.cpp:21-103,131-148); the network arithmetic is the
// reference's conversion/arcface/model_irse.py (see oracle/arcface_oracle.py for the fp32 restatement it is tested against).
//
// Execution plan per forward (IR mode; IR_SE adds gate + apply kernels per unit):
//   stem (direct conv, CUDA cores)  ->  24 units x { [shortcut 1x1 GEMM]  conv1 GEMM (+PReLU)  conv2 GEMM (+bias +shortcut,
//   writes y, BN_next(y), y[::2,::2]) }  ->  folded Linear as split-K GEMM  ->  partial reduce + bias + L2 normalise.
// Every GEMM is conv_gemm_kernel (tcgen05, csrc/conv_kernels.cuh). Activations are fp16 in the shared-halo flat layout.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "common.h"
#include "conv_kernels.cuh"
#include "conv_mt_kernel.cuh"
#include "embed_kernels.cuh"
#include "weights.h"

using namespace frb;

namespace {

constexpr int kGeo[5] = {112, 56, 28, 14, 7};
constexpr int kStageC[5] = {64, 64, 128, 256, 512};   // channels of the map living at geometry g (index 0 = stem output)
constexpr int kStageUnits[5] = {0, 3, 4, 14, 3};
constexpr int kFcSplits = 32;
constexpr int kRunAll = 1 << 30;  // stop_after_unit value for a complete forward
constexpr int kFcK = 64 * 512;  // 8 x 8 padded positions x 512 channels

inline int hpwp(int g) { return (kGeo[g] + 1) * (kGeo[g] + 1); }

struct DevBuf {
    __half* p = nullptr;
    size_t rows = 0;
    int C = 0;
    CUtensorMap tmap{};
};

struct GemmStep {
    CUtensorMap ta{}, tb{};   // tb: weight rows in boxes of `bn`
    CUtensorMap tb64{};       // same weights in boxes of 64 rows: used when 128-wide tiles would leave SMs idle
    bool has64 = false;
    ConvGemmParams prm{};
    int bn = 128;
    int out_geo = 0;     // geometry index of the OUTPUT grid (P = batch * hpwp(out_geo)); -1: FC (P = batch)
    int splits = 1;
};

struct SeStep {
    const __half* u = nullptr;
    const float* fc1 = nullptr;
    const float* fc2 = nullptr;
    int geo = 0, C = 0;
    const __half* res = nullptr;
    int res_mode = 0;
    __half* y = nullptr;
    __half* y_bn = nullptr;
    const float* bn_s = nullptr;
    const float* bn_b = nullptr;
    __half* y_sub = nullptr;
};

struct Step {
    int kind = 0;  // 0 = GEMM, 1 = SE gate+apply
    GemmStep g;
    SeStep se;
    int unit = -1;  // body unit this step belongs to (trace / stop_after)
};

}  // namespace

struct FrEmbedder {
    int device = 0, sms = 0, mode = FR_MODE_IR, max_batch = 0;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;       // everything cudaMalloc'ed (weights + activations)
    // stem
    float *stem_w = nullptr, *stem_b = nullptr, *stem_prelu = nullptr, *u0_bn_s = nullptr, *u0_bn_b = nullptr;
    DevBuf stem_y, stem_yb;
    float* in_f32 = nullptr;         // max_batch x 3 x 112 x 112
    uint8_t* in_u8 = nullptr;        // max_batch x 112 x 112 x 3
    float* gate = nullptr;           // max_batch x 512
    int* se_pool = nullptr;          // SE pooling partials written by conv2: [rows / 32 groups][2 segments][C], largest stage
    float* fc_partial = nullptr;     // kFcSplits x max_batch x 512
    float* fc_bias = nullptr;
    float* out_dev = nullptr;        // max_batch x 512
    std::vector<Step> steps;
    std::vector<std::pair<const __half*, int>> unit_out;  // per unit: (y buffer, geometry index) for fr_embedder_trace
    int last_batch = 0;
    bool last_u8 = false;
    // persistent scratch of fr_embedder_run_boxes (ArcFaceIR50::forward): one frame + its face list, grown on demand
    uint8_t* box_frame = nullptr;
    size_t box_frame_cap = 0;
    void* box_faces = nullptr;
    size_t box_faces_cap = 0;
    GraphCache graphs;
};

namespace {

template <class T>
T* dev_alloc(FrEmbedder* e, size_t count, bool zero) {
    T* p = nullptr;
    FRB_CUDA(cudaMalloc(&p, count * sizeof(T)));
    e->allocs.push_back(p);
    if (zero) FRB_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), e->stream));
    return p;
}

template <class T>
T* upload(FrEmbedder* e, const HostTensor& t) {
    T* p = dev_alloc<T>(e, static_cast<size_t>(t.numel()), false);
    FRB_CUDA(cudaMemcpyAsync(p, t.data, t.nbytes, cudaMemcpyHostToDevice, e->stream));
    return p;
}

DevBuf make_buf(FrEmbedder* e, size_t rows, int C) {
    DevBuf b;
    b.rows = rows;
    b.C = C;
    b.p = dev_alloc<__half>(e, rows * C, true);  // zero once: pad rows/columns are never written afterwards
    b.tmap = make_tmap_2d_f16(b.p, rows, static_cast<uint64_t>(C), 128, 64);
    return b;
}

// Stride-1 3x3 convs run on conv3x3_mt_kernel (persistent, halo-reusing, MT tiles per work unit; conv_mt_kernel.cuh). FR_CONV_MT=0 sends
// them to conv_gemm_kernel instead (A/B); FR_MT_FORCE="bn,mt" pins the tile shape.
const bool g_use_mt = std::getenv("FR_CONV_MT") == nullptr || std::atoi(std::getenv("FR_CONV_MT")) != 0;
int g_conv_sms = 148;  // SM count of the embedder's device (set at create): grid of the persistent conv

constexpr int kSmemBudget = 227 * 1024;

// tap groups + halo geometry of a conv for `tiles` 128-position tiles per CTA (conv_mt_kernel.cuh)
void fill_groups(ConvMtExtra& ex, const ConvGemmParams& prm, int tiles) {
    const int Wp = prm.W + 1;
    if (prm.tap_phase) {
        // stride 2 over a phase-split input: tap (dy, dx) reads phase map (dy != 1, dx != 1) shifted by (dy == 0 ? -1 : 0, dx == 0 ? -1 : 0)
        ex.ngroups = 4;
        ex.halo_chunks = (tiles * kConvBM + Wp + 1 + kConvBM - 1) / kConvBM;
        for (int ph = 0; ph < 4; ++ph) {
            MtGroup& G = ex.grp[ph];
            G.row0 = ph * prm.phase_rows - Wp - 1;
            G.ntaps = 0;
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3, dx = tap % 3;
                if (((dy != 1 ? 2 : 0) + (dx != 1 ? 1 : 0)) != ph) continue;
                G.tap[G.ntaps] = tap;
                G.off[G.ntaps] = Wp + 1 + (dy == 0 ? -Wp : 0) + (dx == 0 ? -1 : 0);
                ++G.ntaps;
            }
        }
    } else {
        ex.ngroups = 1;
        ex.halo_chunks = (tiles * kConvBM + 2 * Wp + 2 + kConvBM - 1) / kConvBM;
        MtGroup& G = ex.grp[0];
        G.row0 = -Wp - 1;
        G.ntaps = 9;
        for (int tap = 0; tap < 9; ++tap) {
            G.tap[tap] = tap;
            G.off[tap] = (tap / 3) * Wp + tap % 3;
        }
    }
}

template <int BN, int MT>
void launch_mt(const GemmStep& s, const ConvGemmParams& prm, cudaStream_t st) {
    ConvMtExtra ex{};
    ex.n_blocks = prm.cout / BN;
    ex.units = ((prm.P + MT * kConvBM - 1) / (MT * kConvBM)) * ex.n_blocks;
    fill_groups(ex, prm, MT);
    const int fixed = conv_mt_smem_bytes(BN, ex.halo_chunks, 0);
    ex.stages = std::min(kMtMaxStages, (kSmemBudget - fixed) / (BN * 128));
    if (ex.stages < 2) throw StateError{"conv3x3_mt_kernel: halo tile does not fit in shared memory"};
    if (!prm.tap_phase && prm.cin_blocks == 1 && ex.n_blocks == 1 && ex.stages >= 9) {
        ex.stationary = 1;
        ex.stages = 9;  // stage index == tap
    }
    const int smem = conv_mt_smem_bytes(BN, ex.halo_chunks, ex.stages);
    const int ctas = std::min(ex.units, g_conv_sms);
    launch_k(conv3x3_mt_kernel<BN, MT>, dim3(ctas), dim3(kMtThreads), smem, st, true, s.ta, BN == 64 && s.bn != 64 ? s.tb64 : s.tb, prm, ex);
}

// CTA-pair version (conv3x3_pair_kernel<BN>): unit = 256 positions x BN channels, weight boxes of BN / 2 rows per CTA
template <int BN>
void launch_pair(const GemmStep& s, const ConvGemmParams& prm, cudaStream_t st) {
    ConvMtExtra ex{};
    ex.n_blocks = prm.cout / BN;
    ex.units = ((prm.P + 2 * kConvBM - 1) / (2 * kConvBM)) * ex.n_blocks;
    fill_groups(ex, prm, 1);
    const int fixed = conv_mt_smem_bytes(BN / 2, ex.halo_chunks, 0);
    ex.stages = std::min(kMtMaxStages, (kSmemBudget - fixed) / ((BN / 2) * 128));
    if (ex.stages < 2) throw StateError{"conv3x3_pair_kernel: halo tile does not fit in shared memory"};
    const int smem = conv_mt_smem_bytes(BN / 2, ex.halo_chunks, ex.stages);
    const int all_pairs = g_conv_sms / 2;
    // a short last round (at most half the pairs busy) is split into half-width items: it then costs half a unit's time
    static const bool split_tail = std::getenv("FR_PAIR_SPLIT") == nullptr || std::atoi(std::getenv("FR_PAIR_SPLIT")) != 0;
    const int rem = ex.units % all_pairs;
    ex.full_units = ex.units;
    ex.half_items = 0;
    if (split_tail && BN == 256 && s.has64 && ex.units > all_pairs && rem > 0 && 2 * rem <= all_pairs) {
        ex.full_units = ex.units - rem;
        ex.half_items = 2 * rem;
    }
    const int pairs = std::min(ex.full_units + ex.half_items, all_pairs);
    launch_k(conv3x3_pair_kernel<BN>, dim3(2 * pairs), dim3(kMtThreads), smem, st, true, s.ta, BN == 128 ? s.tb64 : s.tb, s.tb64, prm, ex);
}

// FR_PAIR=0: never use the CTA-pair kernel; FR_PAIR_BN=128|256 pins its tile width (A/B)
int pick_pair(const GemmStep& s, const ConvGemmParams& prm, double& cost_out) {
    static const bool on = std::getenv("FR_PAIR") == nullptr || std::atoi(std::getenv("FR_PAIR")) != 0;
    static const int force = std::getenv("FR_PAIR_BN") ? std::atoi(std::getenv("FR_PAIR_BN")) : 0;
    if (!on || !g_use_mt || prm.taps != 9 || s.splits != 1 || prm.partial || prm.cout > 512 || s.bn != 128 || !s.has64) return 0;
    const int pairs = g_conv_sms / 2;
    const long long tiles = (prm.P + 2 * kConvBM - 1) / (2 * kConvBM);
    double best = 1e30;
    int best_bn = 0;
    for (int bn : {256, 128}) {
        if (prm.cout % bn || (force && bn != force)) continue;
        const long long units = tiles * (prm.cout / bn);
        // far from filling the GPU: the single-CTA kernel has smaller units. From 3/4 of the pairs on (FR_PAIR_MINFILL, A/B) the pair kernel
        // still wins: each CTA of a pair streams only half of every weight tile from L2 (batch 32: 1.201 -> 1.171 ms, batch 64: 1.767 ->
        // 1.710 ms per forward, profiles/r02_embed_ab_pair_minfill.txt)
        static const double minfill = std::getenv("FR_PAIR_MINFILL") ? std::atof(std::getenv("FR_PAIR_MINFILL")) : 0.75;
        if (units < pairs * minfill) continue;
        // rounds x (unit time + overhead) in units of a single CTA's 128 x 128 tile (pick_mt's scale): a pair does 256 x BN in the time
        // one CTA needs for 128 x BN, faster still because shared memory no longer limits it. Measured (IR-SE-50, batch 256): N = 128
        // MMAs keep the tensor pipe 62-66 % busy; N = 256 pays for a worse last round (225 units on 74 pairs, softened by the half-width
        // split) and still wins, 4.90 vs 5.20 ms per forward
        long long rounds2 = 2 * ((units + pairs - 1) / pairs);  // in half rounds
        const long long rem = units % pairs;
        if (bn == 256 && units > pairs && rem > 0 && 2 * rem <= pairs) rounds2 -= 1;  // short last round split into half-width items
        const double cost = 0.5 * rounds2 * ((bn / 128.0) * (bn == 256 ? 0.7 : 0.85) + 0.5);
        if (cost < best) {
            best = cost;
            best_bn = bn;
        }
    }
    cost_out = best;
    return best_bn;
}

// tile shape of the persistent conv: the largest unit that still gives every SM one; small grids take the smallest
bool pick_mt(const GemmStep& s, const ConvGemmParams& prm, int& bn, int& mt, double& cost_out) {
    cost_out = 1e30;
    if (!g_use_mt || prm.taps != 9 || s.splits != 1 || prm.partial || prm.cout > 512) return false;
    static const bool mt_s2 = std::getenv("FR_MT_S2") == nullptr || std::atoi(std::getenv("FR_MT_S2")) != 0;  // stride-2 convs too (A/B)
    if (prm.tap_phase && !mt_s2) return false;
    static const char* force = std::getenv("FR_MT_FORCE");
    if (force) {
        int fb = 0, fm = 0;
        if (std::sscanf(force, "%d,%d", &fb, &fm) == 2 && (fb == 64 || fb == 128) && (fm == 1 || fm == 2) && prm.cout % fb == 0 &&
            (fb == 64 ? s.has64 || s.bn == 64 : s.bn == 128)) {
            bn = fb;
            mt = fm;
            cost_out = 0.;
            return true;
        }
    }
    // cost model: rounds x (unit work + per-unit overhead), in units of a 128 x 128 tile's MMA time. The overhead (the first halo round
    // trip, the epilogue drain; about half a tile's worth) is what makes small grids prefer ONE round of larger units over two rounds of
    // small ones (measured at batch 32: 1.42 ms with (128,1) forced against 1.59 ms with the smallest units everywhere).
    const int cand[4][2] = {{128, 2}, {128, 1}, {64, 2}, {64, 1}};
    double best = 1e30;
    bn = 0;
    for (const auto& c : cand) {
        if (c[0] == 128 && (s.bn != 128 || prm.cout % 128)) continue;
        if (c[0] == 64 && !(s.bn == 64 || s.has64)) continue;
        const long long units = static_cast<long long>((prm.P + c[1] * kConvBM - 1) / (c[1] * kConvBM)) * (prm.cout / c[0]);
        const long long rounds = (units + g_conv_sms - 1) / g_conv_sms;
        const double work = c[1] * (c[0] / 128.0) * (c[0] == 64 ? 1.25 : 1.0);  // N = 64 MMAs are bound by the A-operand read
        const double cost = static_cast<double>(rounds) * (work + 0.5);
        if (cost < best - 1e-9) {
            best = cost;
            bn = c[0];
            mt = c[1];
        }
    }
    cost_out = best;
    return bn != 0;
}

template <int BN>
void launch_gemm(const GemmStep& s, int P, cudaStream_t st) {
    ConvGemmParams prm = s.prm;
    prm.P = P;
    dim3 grid((P + kConvBM - 1) / kConvBM, prm.cout / BN, s.splits);
    if (prm.pool) launch_k(conv_gemm_kernel<BN, true, false>, grid, dim3(kConvThreads), ConvCfg<BN>::kSmemBytes, st, true, s.ta, s.tb, prm);
    else launch_k(conv_gemm_kernel<BN, false, false>, grid, dim3(kConvThreads), ConvCfg<BN>::kSmemBytes, st, true, s.ta, s.tb, prm);
    count_launch();
}

void launch_conv(const GemmStep& g0, int P, int sms, cudaStream_t st) {
    GemmStep g = g0;
    g.prm.P = P;
    int bn = 0, mt = 0;
    double pair_cost = 1e30, mt_cost = 1e30;
    const int pbn = pick_pair(g, g.prm, pair_cost);
    const bool mt_ok = pick_mt(g, g.prm, bn, mt, mt_cost);
    if (pbn && (!mt_ok || pair_cost <= mt_cost)) {
        if (pbn == 256) launch_pair<256>(g, g.prm, st);
        else launch_pair<128>(g, g.prm, st);
        count_launch();
        return;
    }
    if (mt_ok) {
        if (bn == 128 && mt == 2) launch_mt<128, 2>(g, g.prm, st);
        else if (bn == 128) launch_mt<128, 1>(g, g.prm, st);
        else if (mt == 2) launch_mt<64, 2>(g, g.prm, st);
        else launch_mt<64, 1>(g, g.prm, st);
        count_launch();
        return;
    }
    // 128-wide channel tiles only when they still give every SM two CTAs' worth of work; otherwise 64-wide tiles (two CTAs
    // fit per SM, so one CTA's epilogue overlaps another's main loop)
    const long long tiles128 = static_cast<long long>((P + kConvBM - 1) / kConvBM) * (g.prm.cout / 128);
    if (g.bn == 128 && g.has64 && tiles128 < 2LL * sms) {
        g.bn = 64;
        g.tb = g.tb64;
    }
    if (g.bn == 64) launch_gemm<64>(g, P, st);
    else launch_gemm<128>(g, P, st);
}

void run_steps(FrEmbedder* e, int batch, bool u8_input, int stop_after_unit, cudaStream_t st = nullptr) {
    if (!st) st = e->stream;
    {   // tensor-core stem (arcface_stem_tc_kernel): persistent, a few CTAs per SM
        const int tiles = (batch * hpwp(0) + 127) / 128;
        const int ctas = std::min(tiles, e->sms * 6);
        if (u8_input)
            launch_k(arcface_stem_tc_kernel<true>, dim3(ctas), dim3(128), 0, st, true, static_cast<const void*>(e->in_u8), batch, e->stem_w,
                     e->stem_b, e->stem_prelu, e->u0_bn_s, e->u0_bn_b, e->stem_y.p, e->stem_yb.p);
        else
            launch_k(arcface_stem_tc_kernel<false>, dim3(ctas), dim3(128), 0, st, true, static_cast<const void*>(e->in_f32), batch, e->stem_w,
                     e->stem_b, e->stem_prelu, e->u0_bn_s, e->u0_bn_b, e->stem_y.p, e->stem_yb.p);
    }
    count_launch();
    for (const Step& s : e->steps) {
        if (s.unit > stop_after_unit) break;
        if (s.kind == 0) {
            const int P = s.g.out_geo >= 0 ? batch * hpwp(s.g.out_geo) : batch;
            GemmStep g = s.g;
            if (g.out_geo < 0) {  // FC: one GEMM row per image
                g.prm.W = batch;
                g.prm.H = 1;
            }
            launch_conv(g, P, e->sms, st);
        } else {
            const SeStep& q = s.se;
            const int H = kGeo[q.geo], P = batch * hpwp(q.geo);
            launch_k(se_gate_kernel, dim3(batch), dim3(512), 0, st, true, e->se_pool, H, H, q.C, q.fc1, q.fc2, e->gate);
            const long long items = static_cast<long long>(P) * (q.C / 8);
            const int blocks = static_cast<int>(std::min<long long>((items + 511) / 512, 8LL * e->sms));  // 2 items per thread per pass
            launch_k(se_apply_kernel, dim3(blocks), dim3(256), 0, st, true, q.u, e->gate, P, H, H, q.C, q.res, q.res_mode, q.y, q.y_bn, q.bn_s,
                     q.bn_b, q.y_sub);
            count_launch();
            count_launch();
        }
    }
    if (stop_after_unit == kRunAll) {
        launch_k(fc_reduce_l2norm_kernel, dim3(batch), dim3(512), 0, st, true, e->fc_partial, kFcSplits, batch, e->fc_bias, e->out_dev);
        count_launch();
    }
    FRB_CUDA(cudaGetLastError());
}

// complete forward, replayed from a CUDA graph per (batch, input kind)
void forward_all(FrEmbedder* e, int batch, bool u8_input, cudaStream_t st = nullptr) {
    NvtxRange nvtx("fr.embed.forward");
    if (!st) st = e->stream;
    e->graphs.run({static_cast<uint64_t>(batch), u8_input ? 1ull : 0ull, 0ull}, st, [&] { run_steps(e, batch, u8_input, kRunAll, st); });
}

void build_plan(FrEmbedder* e, const WeightFile& wf) {
    const bool se = e->mode == FR_MODE_IR_SE;
    const int B = e->max_batch;
    auto f32 = [&](const std::string& n, int64_t numel) { return upload<float>(e, wf.get(n, 0, numel)); };
    auto f16 = [&](const std::string& n, int64_t numel) { return upload<__half>(e, wf.get(n, 1, numel)); };

    e->stem_w = f32("stem.w", 64 * 27);
    e->stem_b = f32("stem.b", 64);
    e->stem_prelu = f32("stem.prelu", 64);
    e->in_f32 = dev_alloc<float>(e, static_cast<size_t>(B) * 3 * 112 * 112, false);
    e->in_u8 = dev_alloc<uint8_t>(e, static_cast<size_t>(B) * 112 * 112 * 3, false);
    if (se) {
        e->gate = dev_alloc<float>(e, static_cast<size_t>(B) * 512, false);
        size_t pool_floats = 0;
        for (int g = 1; g <= 4; ++g)
            pool_floats = std::max(pool_floats, (static_cast<size_t>(B) * hpwp(g) / 32 + 8) * 2 * kStageC[g]);
        e->se_pool = dev_alloc<int>(e, pool_floats, true);
    }
    e->fc_partial = dev_alloc<float>(e, static_cast<size_t>(kFcSplits) * B * 512, false);
    e->out_dev = dev_alloc<float>(e, static_cast<size_t>(B) * 512, false);
    e->stem_y = make_buf(e, static_cast<size_t>(B) * hpwp(0), 64);
    e->stem_yb = make_buf(e, static_cast<size_t>(B) * hpwp(0), 64);

    // per-unit parameters
    struct UnitW {
        int cin, d, stride;
        float *bn1_s, *bn1_b, *prelu, *b2, *bs = nullptr, *fc1 = nullptr, *fc2 = nullptr;
        __half *w1, *w2, *ws = nullptr;
        CUtensorMap t1, t2, ts, t1n, t2n, tsn;  // *n: 64-row boxes
    };
    std::vector<UnitW> uw;
    {
        int cin = 64;
        for (int stage = 1; stage <= 4; ++stage) {
            for (int j = 0; j < kStageUnits[stage]; ++j) {
                UnitW u{};
                u.cin = (j == 0) ? cin : kStageC[stage];
                u.d = kStageC[stage];
                u.stride = (j == 0) ? 2 : 1;
                const std::string q = "u" + std::to_string(uw.size()) + ".";
                u.bn1_s = f32(q + "bn1.s", u.cin);
                u.bn1_b = f32(q + "bn1.b", u.cin);
                u.w1 = f16(q + "conv1.w", static_cast<int64_t>(u.d) * 9 * u.cin);
                u.prelu = f32(q + "prelu", u.d);
                u.w2 = f16(q + "conv2.w", static_cast<int64_t>(u.d) * 9 * u.d);
                u.b2 = f32(q + "conv2.b", u.d);
                const int bn = u.d == 64 ? 64 : 128;
                u.t1 = make_tmap_2d_f16(u.w1, u.d, 9ull * u.cin, bn, 64);
                u.t2 = make_tmap_2d_f16(u.w2, u.d, 9ull * u.d, bn, 64);
                u.t1n = make_tmap_2d_f16(u.w1, u.d, 9ull * u.cin, 64, 64);
                u.t2n = make_tmap_2d_f16(u.w2, u.d, 9ull * u.d, 64, 64);
                if (u.cin != u.d) {
                    u.ws = f16(q + "sc.w", static_cast<int64_t>(u.d) * u.cin);
                    u.bs = f32(q + "sc.b", u.d);
                    u.ts = make_tmap_2d_f16(u.ws, u.d, u.cin, bn, 64);
                    u.tsn = make_tmap_2d_f16(u.ws, u.d, u.cin, 64, 64);
                }
                if (se) {
                    u.fc1 = f32(q + "se.fc1", static_cast<int64_t>(u.d / 16) * u.d);
                    u.fc2 = f32(q + "se.fc2", static_cast<int64_t>(u.d) * (u.d / 16));
                }
                uw.push_back(u);
            }
            cin = kStageC[stage];
        }
    }
    e->u0_bn_s = uw[0].bn1_s;
    e->u0_bn_b = uw[0].bn1_b;

    // activation buffers per stage
    struct StageBufs {
        DevBuf tphase, t, y[2], yb[2], sc, ysub, u;
    };
    StageBufs sb[5];
    for (int s = 1; s <= 4; ++s) {
        const size_t P = static_cast<size_t>(B) * hpwp(s);
        const int C = kStageC[s];
        sb[s].tphase = make_buf(e, 4 * P, C);
        sb[s].t = make_buf(e, P, C);
        for (int k = 0; k < 2; ++k) {
            sb[s].y[k] = make_buf(e, P, C);
            sb[s].yb[k] = make_buf(e, P, C);
        }
        if (s >= 2) {
            sb[s].sc = make_buf(e, P, C);
            sb[s].ysub = make_buf(e, P, kStageC[s - 1]);
        }
        if (se) sb[s].u = make_buf(e, P, C);
    }

    auto base_prm = [](int geo, int cin, int taps, int cout) {
        ConvGemmParams p{};
        p.H = p.W = kGeo[geo];
        p.cin_blocks = cin / 64;
        p.taps = taps;
        p.cout = cout;
        p.kb_per_split = taps * (cin / 64);
        return p;
    };

    int ui = 0;
    const DevBuf* in_y = &e->stem_y;
    const DevBuf* in_yb = &e->stem_yb;
    for (int stage = 1; stage <= 4; ++stage) {
        int cur = 0;
        for (int j = 0; j < kStageUnits[stage]; ++j, ++ui) {
            const UnitW& u = uw[ui];
            const bool first = j == 0;
            const bool last_unit = ui + 1 == static_cast<int>(uw.size());
            const int bn = u.d == 64 ? 64 : 128;
            const int in_geo = first ? stage - 1 : stage;
            // ---- shortcut 1x1 stride-2 conv + BN (folded) on the subsampled input (model_irse.py:54-56)
            if (first && u.cin != u.d) {
                Step s;
                s.unit = ui;
                s.g.ta = sb[stage].ysub.tmap;
                s.g.tb = u.ts;
                s.g.tb64 = u.tsn;
                s.g.has64 = true;
                s.g.bn = bn;
                s.g.out_geo = stage;
                s.g.prm = base_prm(stage, u.cin, 1, u.d);
                s.g.prm.bias = u.bs;
                s.g.prm.out = sb[stage].sc.p;
                e->steps.push_back(s);
            }
            // ---- conv1: 3x3 stride 1 on BN1(x) + PReLU (:57-59); a stride-2 unit stores it phase-split for conv2
            {
                Step s;
                s.unit = ui;
                s.g.ta = in_yb->tmap;
                s.g.tb = u.t1;
                s.g.tb64 = u.t1n;
                s.g.has64 = true;
                s.g.bn = bn;
                s.g.out_geo = in_geo;
                s.g.prm = base_prm(in_geo, u.cin, 9, u.d);
                s.g.prm.prelu = u.prelu;
                if (first) {
                    s.g.prm.out = sb[stage].tphase.p;
                    s.g.prm.out_mode = kOutPhaseSplit;
                    s.g.prm.out_phase_rows = static_cast<long long>(B) * hpwp(stage);
                } else {
                    s.g.prm.out = sb[stage].t.p;
                }
                e->steps.push_back(s);
            }
            // ---- conv2: 3x3 stride s + BN (folded) [+ shortcut] (:59-65)
            const int nxt = first ? 0 : cur ^ 1;
            const __half* res = nullptr;
            int res_mode = kResNone;
            if (first && u.cin == u.d) {
                res = in_y->p;  // MaxPool2d(1, 2): x[::2, ::2]
                res_mode = kResSubsample;
            } else if (first) {
                res = sb[stage].sc.p;
                res_mode = kResSame;
            } else {
                res = in_y->p;
                res_mode = kResSame;
            }
            __half* y = sb[stage].y[nxt].p;
            __half* yb = last_unit ? nullptr : sb[stage].yb[nxt].p;
            const float* nbs = last_unit ? nullptr : uw[ui + 1].bn1_s;
            const float* nbb = last_unit ? nullptr : uw[ui + 1].bn1_b;
            const bool next_is_first = !last_unit && (j + 1 == kStageUnits[stage]);
            __half* ysub = next_is_first ? sb[stage + 1].ysub.p : nullptr;
            {
                Step s;
                s.unit = ui;
                s.g.ta = first ? sb[stage].tphase.tmap : sb[stage].t.tmap;
                s.g.tb = u.t2;
                s.g.tb64 = u.t2n;
                s.g.has64 = true;
                s.g.bn = bn;
                s.g.out_geo = stage;
                s.g.prm = base_prm(stage, u.d, 9, u.d);
                s.g.prm.bias = u.b2;
                if (first) {
                    s.g.prm.tap_phase = 1;
                    s.g.prm.phase_rows = B * hpwp(stage);
                }
                if (se) {
                    s.g.prm.out = sb[stage].u.p;
                    s.g.prm.pool = e->se_pool;
                } else {
                    s.g.prm.res = res;
                    s.g.prm.res_mode = res_mode;
                    s.g.prm.out = y;
                    s.g.prm.out_bn = yb;
                    s.g.prm.bn_s = nbs;
                    s.g.prm.bn_b = nbb;
                    s.g.prm.out_sub = ysub;
                }
                e->steps.push_back(s);
            }
            if (se) {
                Step s;
                s.kind = 1;
                s.unit = ui;
                s.se.u = sb[stage].u.p;
                s.se.fc1 = u.fc1;
                s.se.fc2 = u.fc2;
                s.se.geo = stage;
                s.se.C = u.d;
                s.se.res = res;
                s.se.res_mode = res_mode;
                s.se.y = y;
                s.se.y_bn = yb;
                s.se.bn_s = nbs;
                s.se.bn_b = nbb;
                s.se.y_sub = ysub;
                e->steps.push_back(s);
            }
            e->unit_out.emplace_back(y, stage);
            in_y = &sb[stage].y[nxt];
            in_yb = &sb[stage].yb[nxt];
            cur = nxt;
        }
    }
    // ---- output_layer: BN2d + Flatten + Linear + BN1d folded into one GEMM over the padded 8x8x512 map (:142-147)
    {
        __half* fcw = f16("fc.w", 512ll * kFcK);
        e->fc_bias = f32("fc.b", 512);
        Step s;
        s.unit = 1000;
        s.g.ta = make_tmap_2d_f16(in_y->p, static_cast<uint64_t>(B), kFcK, 128, 64);
        s.g.tb = make_tmap_2d_f16(fcw, 512, kFcK, 128, 64);
        s.g.bn = 128;
        s.g.out_geo = -1;
        s.g.splits = kFcSplits;
        s.g.prm.cin_blocks = kFcK / 64;
        s.g.prm.taps = 1;
        s.g.prm.cout = 512;
        s.g.prm.kb_per_split = (kFcK / 64) / kFcSplits;
        s.g.prm.partial = e->fc_partial;
        e->steps.push_back(s);
    }
    FRB_CUDA(cudaStreamSynchronize(e->stream));
}

void check_batch(const FrEmbedder* e, const void* in, int batch, const void* out) {
    if (!e) throw ArgError{"null embedder"};
    if (!in || !out) throw ArgError{"null buffer"};
    if (batch < 1 || batch > e->max_batch) throw ArgError{"batch out of range (1..max_batch)"};
}

}  // namespace

// internal hooks for the end-to-end pipeline (csrc/pipeline.cu): crops are written straight into the embedder's u8 input
namespace frb {
uint8_t* embedder_u8_input(FrEmbedder* e) { return e->in_u8; }
float* embedder_output(FrEmbedder* e) { return e->out_dev; }
int embedder_max_batch(const FrEmbedder* e) { return e->max_batch; }
int embedder_device(const FrEmbedder* e) { return e->device; }
cudaStream_t embedder_stream(FrEmbedder* e) { return e->stream; }
// grow-only device scratch owned by the embedder (freed with it); the caller has synchronised the embedder's stream
static void* grow_scratch(FrEmbedder* e, void*& buf, size_t& cap, size_t bytes) {
    if (bytes <= cap) return buf;
    void* nb = nullptr;
    FRB_CUDA(cudaMalloc(&nb, bytes));
    if (buf) {
        cudaFree(buf);
        for (auto& a : e->allocs)
            if (a == buf) a = nb;
    } else {
        e->allocs.push_back(nb);
    }
    buf = nb;
    cap = bytes;
    return buf;
}
uint8_t* embedder_frame_scratch(FrEmbedder* e, size_t bytes) {
    void* b = e->box_frame;
    grow_scratch(e, b, e->box_frame_cap, bytes);
    e->box_frame = static_cast<uint8_t*>(b);
    return e->box_frame;
}
void* embedder_faces_scratch(FrEmbedder* e, size_t bytes) { return grow_scratch(e, e->box_faces, e->box_faces_cap, bytes); }
void embedder_forward_u8(FrEmbedder* e, int batch, cudaStream_t st) {
    e->last_batch = batch;
    e->last_u8 = true;
    forward_all(e, batch, true, st);
}
}  // namespace frb

extern "C" {

int fr_embedder_create(const char* weights_path, int max_batch, int device, FrEmbedder** out) {
    return guarded([&] {
        if (!out) throw ArgError{"out is null"};
        if (max_batch < 1 || max_batch > 1024) throw ArgError{"max_batch out of range (1..1024)"};
        WeightFile wf = load_weight_file(weights_path);
        if (wf.kind != kKindArcfaceIR && wf.kind != kKindArcfaceIRSE) throw FileError{FR_EFORMAT, "weight file is not an ArcFace checkpoint"};
        std::unique_ptr<FrEmbedder> e(new FrEmbedder());
        e->sms = use_device(device);
        e->device = device;
        e->mode = wf.kind == kKindArcfaceIRSE ? FR_MODE_IR_SE : FR_MODE_IR;
        e->max_batch = max_batch;
        try {
            FRB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
            FRB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64>::kSmemBytes));
            FRB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<128>::kSmemBytes));
            FRB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64>::kSmemBytes));
            FRB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<128>::kSmemBytes));
            g_conv_sms = e->sms;
            FRB_CUDA(cudaFuncSetAttribute(conv3x3_mt_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
            FRB_CUDA(cudaFuncSetAttribute(conv3x3_mt_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
            FRB_CUDA(cudaFuncSetAttribute(conv3x3_mt_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
            FRB_CUDA(cudaFuncSetAttribute(conv3x3_mt_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
            FRB_CUDA(cudaFuncSetAttribute(conv3x3_pair_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
            FRB_CUDA(cudaFuncSetAttribute(conv3x3_pair_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
            build_plan(e.get(), wf);
        } catch (...) {
            fr_embedder_destroy(e.release());
            throw;
        }
        *out = e.release();
    });
}

void fr_embedder_destroy(FrEmbedder* e) {
    if (!e) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (void* p : e->allocs) cudaFree(p);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (prev >= 0) cudaSetDevice(prev);
    delete e;
}

int fr_embedder_mode(const FrEmbedder* e) { return e ? e->mode : FR_EINVAL; }
int fr_embedder_max_batch(const FrEmbedder* e) { return e ? e->max_batch : FR_EINVAL; }

int fr_embedder_run(FrEmbedder* e, const float* chw, int batch, float* out512) {
    return guarded([&] {
        check_batch(e, chw, batch, out512);
        DeviceGuard dg(e->device);
        FRB_CUDA(cudaMemcpyAsync(e->in_f32, chw, sizeof(float) * batch * 3 * 112 * 112, cudaMemcpyHostToDevice, e->stream));
        e->last_batch = batch;
        e->last_u8 = false;
        forward_all(e, batch, false);
        FRB_CUDA(cudaMemcpyAsync(out512, e->out_dev, sizeof(float) * batch * 512, cudaMemcpyDeviceToHost, e->stream));
        FRB_CUDA(cudaStreamSynchronize(e->stream));
    });
}

int fr_embedder_run_crops(FrEmbedder* e, const uint8_t* crops_bgr_u8, int batch, float* out512) {
    return guarded([&] {
        check_batch(e, crops_bgr_u8, batch, out512);
        DeviceGuard dg(e->device);
        FRB_CUDA(cudaMemcpyAsync(e->in_u8, crops_bgr_u8, static_cast<size_t>(batch) * 112 * 112 * 3, cudaMemcpyHostToDevice, e->stream));
        e->last_batch = batch;
        e->last_u8 = true;
        forward_all(e, batch, true);
        FRB_CUDA(cudaMemcpyAsync(out512, e->out_dev, sizeof(float) * batch * 512, cudaMemcpyDeviceToHost, e->stream));
        FRB_CUDA(cudaStreamSynchronize(e->stream));
    });
}

int fr_embedder_run_dev(FrEmbedder* e, const float* chw_dev, int batch, float* out512_dev, void* stream) {
    return guarded([&] {
        check_batch(e, chw_dev, batch, out512_dev);
        DeviceGuard dg(e->device);
        // the plan runs on the handle's stream; order it after / before the caller's stream with events
        // (stream == NULL is the legacy default stream: the handle's stream is non-blocking and does not synchronise with it implicitly)
        cudaStream_t user = static_cast<cudaStream_t>(stream);
        cudaEvent_t ev;
        FRB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        FRB_CUDA(cudaEventRecord(ev, user));
        FRB_CUDA(cudaStreamWaitEvent(e->stream, ev, 0));
        FRB_CUDA(cudaMemcpyAsync(e->in_f32, chw_dev, sizeof(float) * batch * 3 * 112 * 112, cudaMemcpyDeviceToDevice, e->stream));
        e->last_batch = batch;
        e->last_u8 = false;
        forward_all(e, batch, false);
        FRB_CUDA(cudaMemcpyAsync(out512_dev, e->out_dev, sizeof(float) * batch * 512, cudaMemcpyDeviceToDevice, e->stream));
        FRB_CUDA(cudaEventRecord(ev, e->stream));
        FRB_CUDA(cudaStreamWaitEvent(user, ev, 0));
        cudaEventDestroy(ev);
    });
}

int fr_embedder_trace(FrEmbedder* e, int layer, float* out, int64_t cap, int64_t* n_written) {
    return guarded([&] {
        if (!e || !out || !n_written) throw ArgError{"null argument"};
        if (e->last_batch < 1) throw StateError{"fr_embedder_trace: no previous run"};
        if (layer < 0 || layer > static_cast<int>(e->unit_out.size())) throw ArgError{"layer out of range (0 = input layer, 1..24 = units)"};
        DeviceGuard dg(e->device);
        const int batch = e->last_batch;
        const __half* src;
        int geo;
        if (layer == 0) {
            src = e->stem_y.p;
            geo = 0;
        } else {
            src = e->unit_out[layer - 1].first;
            geo = e->unit_out[layer - 1].second;
        }
        const int H = kGeo[geo], C = kStageC[geo];
        const int64_t total = static_cast<int64_t>(batch) * C * H * H;
        if (cap < total) throw ArgError{"trace buffer too small"};
        run_steps(e, batch, e->last_u8, layer - 1);  // re-run the retained input up to that unit (later units reuse buffers)
        float* tmp = nullptr;
        FRB_CUDA(cudaMalloc(&tmp, sizeof(float) * total));
        unpack_nchw_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, e->stream>>>(src, batch, H, H, C, tmp);
        count_launch();
        cudaError_t err = cudaMemcpyAsync(out, tmp, sizeof(float) * total, cudaMemcpyDeviceToHost, e->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
        cudaFree(tmp);
        FRB_CUDA(err);
        *n_written = total;
    });
}

}  // extern "C"
