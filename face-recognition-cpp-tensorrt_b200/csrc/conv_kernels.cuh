// Implicit-GEMM convolution on tcgen05 tensor cores for the ArcFace IR-(SE)50 body and the RetinaFace 64-channel convs
// (network spec: /root/reference conversion/arcface/model_irse.py:48-90,139-147; conversion/retina/models/net.py).
//
// Activation layout ("shared-halo flat NHWC"): a feature map of B images of H x W pixels with C channels (fp16) is a matrix
// [P, C], P = B * (H+1) * (W+1): position p = img*(H+1)*(W+1) + r*(W+1) + c. Column c == W and row r == H of every image
// are zero and are never written, so the 3x3 neighbour (dy, dx) of position p is simply row p + (dy-1)*(W+1) + (dx-1) of the
// matrix — the zero column / row (and TMA's zero fill before row 0 and after row P-1) supply the convolution's zero padding.
// A conv is then  D[p, n] = sum_tap sum_c  X[p + shift(tap), c] * Wt[n, tap*Cin + c] : for each tap a plain 2-D TMA box
// [128 positions x 64 channels] at a shifted row coordinate, no im2col buffer.
// Stride-2 3x3 convs read a "phase-split" input written by their producer: four (H/2 x W/2) maps X_pq[i,j] = X[2i+p, 2j+q] in
// the same layout; tap (dy,dx) reads phase (dy!=1, dx!=1) shifted by (dy==0 ? -1 : 0, dx==0 ? -1 : 0).
//
// One CTA computes a [128 positions x BN output channels] tile: warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers -> bias / PReLU / residual / next-unit BN -> fp16 stores).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx_sm100.cuh"

namespace frb {

constexpr int kConvBM = 128;
constexpr int kConvThreads = 256;
// pipeline depth per tile width: two CTAs must fit per SM so that one CTA's epilogue overlaps the other's main loop
// (BN = 128: 3 x 32 KiB, BN <= 64: 4 x <= 24 KiB)
template <int BN>
struct ConvStages {
    static constexpr int value = BN >= 128 ? 3 : 4;
};

enum ConvOutMode { kOutNormal = 0, kOutPhaseSplit = 1 };
enum ConvResMode { kResNone = 0, kResSame = 1, kResSubsample = 2, kResUpsample = 3 };

struct ConvGemmParams {
    int P;             // output positions (GEMM rows)
    int H, W;          // valid extent of the output geometry (Wp = W+1, HpWp = (H+1)*(W+1))
    int cin_blocks;    // Cin / 64
    int taps;          // 9 or 1
    int tap_phase;     // 1: stride-2 taps over a phase-split input of the output's geometry
    int phase_rows;    // rows per phase map (tap_phase) = P
    int cout;          // total output channels (row stride of the outputs)
    int kb_per_split;  // k-blocks per blockIdx.z (split-K); taps*cin_blocks when not split
    const float* bias;     // [cout] or null
    const float* prelu;    // [cout] or null : PReLU slopes applied after bias
    const __half* res;     // residual source or null
    int res_mode;          // ConvResMode; kResSubsample: source geometry (2H, 2W), read at (2r, 2c);
                           // kResUpsample: source geometry (H/2, W/2), read at (r/2, c/2) (nearest x2, FPN top-down add)
    int relu;              // ReLU after bias (before the residual add)
    int ld_out;            // row stride (channels) of out / out_bn / out_sub; 0 = cout. out may point into a channel slice
    int ld_res;            // row stride of res; 0 = cout
    __half* out;           // [P_out, cout] fp16 or null
    int out_mode;          // ConvOutMode; kOutPhaseSplit: out geometry (H/2, W/2) x 4 phases
    long long out_phase_rows;  // rows between the phase maps of `out` (fixed at the handle's maximum batch)
    __half* out_bn;        // second output: out * bn_s + bn_b (the next unit's pre-activation BatchNorm), same layout as out
    const float* bn_s;
    const float* bn_b;
    __half* out_sub;       // third output: out at even (r, c) written in geometry (H/2, W/2) (input of a stride-2 1x1 shortcut)
    float* partial;        // split-K: fp32 accumulators [split][P][cout], no epilogue math
    __half* out2;          // channels >= out2_from go to out2 (row stride ld_out2, channel index rebased) instead of out: one GEMM
    int out2_from, ld_out2;  // feeding two destination maps (RetinaFace SSH: conv3X3 and conv5X5_1 share their input)
    // HEADS epilogue (conv_gemm_kernel<32, false, true>): the 32 GEMM columns are the three 1x1 heads of a RetinaFace level
    float* head_loc;       // [batch][anchors_total][4]
    float* head_conf;      // [batch][anchors_total][2], softmax applied
    float* head_landm;     // [batch][anchors_total][10] or null
    int anchors_total, level_offset;
    int* pool;             // SE pooling partials or null: [P/32 groups][2 segments][cout] exact fixed-point channel sums of the conv output (after
                           // bias / PReLU / ReLU, before the residual), see pool_store16
};

template <int BN>
struct ConvCfg {
    static constexpr int kConvStages = ConvStages<BN>::value;
    static constexpr int kABytes = kConvBM * 128;
    static constexpr int kBBytes = BN * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kParamBytes = 4 * BN * 4;  // bias, prelu, bn_s, bn_b of this CTA's channel slice (fp32)
    static constexpr int kSmemBytes = 1024 + kConvStages * kStageBytes + 256 + kParamBytes;
    static constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : 128);
};

// SE pooling partials are EXACT: every conv output is converted to fixed point (kPoolScale = 2^14, clamped to +-2^26, i.e. |v| < 4096
// with a resolution of 6.1e-5 - finer than the fp16 the map itself is stored in) before it is summed, so the per-image channel sums do
// not depend on the order of summation: an image's embedding is bit-identical wherever it sits in a batch, although the 32-position
// groups are aligned to the batch's matrix rows, not to the image. 32 clamped values cannot overflow an int32.
constexpr float kPoolScale = 16384.f;  // 2^14
constexpr int kPoolClamp = (1 << 26) - 1;
typedef int pool_t;
__device__ __forceinline__ pool_t pool_fix(float v) { return max(-kPoolClamp, min(kPoolClamp, __float2int_rn(v * kPoolScale))); }

// Sum of v[j] over the 32 lanes of a warp for 16 values at once: 16 shuffles instead of 80. On return every lane holds the total of
// value index (lane >> 1).
__device__ __forceinline__ pool_t warp_sum16(const pool_t (&v)[16], int lane) {
    constexpr unsigned kFull = 0xffffffffu;
    pool_t w8[8], w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const pool_t send = b4 ? v[j] : v[j + 8], keep = b4 ? v[j + 8] : v[j];
        w8[j] = keep + __shfl_xor_sync(kFull, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const pool_t send = b3 ? w8[j] : w8[j + 4], keep = b3 ? w8[j + 4] : w8[j];
        w4[j] = keep + __shfl_xor_sync(kFull, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const pool_t send = b2 ? w4[j] : w4[j + 2], keep = b2 ? w4[j + 2] : w4[j];
        w2[j] = keep + __shfl_xor_sync(kFull, send, 4);
    }
    const pool_t send = b1 ? w2[0] : w2[1], keep = b1 ? w2[1] : w2[0];
    pool_t w1 = keep + __shfl_xor_sync(kFull, send, 2);
    w1 += __shfl_xor_sync(kFull, w1, 1);
    return w1;
}

// SE pooling partials of one warp's 32 positions x 16 channels (global channel n..n+15): pool[(group * 2 + seg) * cout + channel],
// group = position / 32, seg 0 = positions of the image the group's first position lies in, seg 1 = positions of the next image
// (an image has >= 64 positions, so a group touches at most two). v: the conv outputs of this lane's position; valid = the position
// is a pixel (pads and rows beyond P count as zero).
__device__ __forceinline__ void pool_store16(pool_t* __restrict__ pool, int cout, int group, int n, const float (&v)[16], bool valid, int lane,
                                             bool straddles, bool in_first_image) {
    pool_t q[16], a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) q[j] = valid ? pool_fix(v[j]) : 0;
    if (!straddles) {  // warp-uniform
        const pool_t s0 = warp_sum16(q, lane);
        if (!(lane & 1)) pool[(static_cast<size_t>(group) * 2) * cout + n + (lane >> 1)] = s0;
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = in_first_image ? q[j] : 0;
        const pool_t s0 = warp_sum16(a, lane);
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = in_first_image ? 0 : q[j];
        const pool_t s1 = warp_sum16(a, lane);
        if (!(lane & 1)) {
            pool[(static_cast<size_t>(group) * 2) * cout + n + (lane >> 1)] = s0;
            pool[(static_cast<size_t>(group) * 2 + 1) * cout + n + (lane >> 1)] = s1;
        }
    }
}

// POOL: also write the SE pooling partials (prm.pool); a separate instantiation because the shuffle tree costs registers that would
// push the plain kernel below two CTAs per SM.
template <int BN, bool POOL = false, bool HEADS = false>
__global__ void __launch_bounds__(kConvThreads, POOL ? 1 : 2)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ ConvGemmParams prm) {
    using Cfg = ConvCfg<BN>;
    constexpr int kConvStages = Cfg::kConvStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kConvStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + kConvStages;
    uint64_t* acc_bar = empty_bar + kConvStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    float* s_bias = reinterpret_cast<float*>(smem + kConvStages * Cfg::kStageBytes + 256);
    float* s_prelu = s_bias + BN;
    float* s_bns = s_prelu + BN;
    float* s_bnb = s_bns + BN;

    griddep_launch_dependents();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int p0 = blockIdx.x * kConvBM;
    const int n0 = blockIdx.y * BN;
    const int Wp = prm.W + 1;
    const int kb_begin = blockIdx.z * prm.kb_per_split;
    const int kb_total = prm.taps * prm.cin_blocks;
    const int kb_end = min(kb_total, kb_begin + prm.kb_per_split);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kConvStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    if (warp >= 4) {  // epilogue parameters of this CTA's channel slice -> shared memory (the epilogue never touches global for them)
        for (int i = threadIdx.x - 128; i < BN; i += 128) {
            const int n = blockIdx.y * BN + i;
            s_bias[i] = prm.bias ? __ldg(prm.bias + n) : 0.f;
            s_prelu[i] = prm.prelu ? __ldg(prm.prelu + n) : 1.f;
            s_bns[i] = prm.out_bn ? __ldg(prm.bn_s + n) : 1.f;
            s_bnb[i] = prm.out_bn ? __ldg(prm.bn_b + n) : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();  // everything above touched only this CTA's state and static parameters

    if (warp == 0) {
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                const int tap = kb / prm.cin_blocks;
                const int cb = kb - tap * prm.cin_blocks;
                int row;
                if (prm.taps == 1) {
                    row = p0;
                } else {
                    const int dy = tap / 3, dx = tap - dy * 3;
                    if (prm.tap_phase) {
                        const int ph = (dy != 1 ? 2 : 0) + (dx != 1 ? 1 : 0);
                        row = ph * prm.phase_rows + p0 + (dy == 0 ? -Wp : 0) + (dx == 0 ? -1 : 0);
                    } else {
                        row = p0 + (dy - 1) * Wp + (dx - 1);
                    }
                }
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* a_dst = smem + stage * Cfg::kStageBytes;
                uint8_t* b_dst = a_dst + Cfg::kABytes;
                mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                tma_load_2d(a_dst, &tmap_a, &full_bar[stage], cb * 64, row, kEvictNormal);
                tma_load_2d(b_dst, &tmap_b, &full_bar[stage], kb * 64, n0, kEvictLast);
                if (++stage == kConvStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(kConvBM, BN, 0, 0);
            uint32_t stage = 0, phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16_ss(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                (kb > kb_begin || k > 0) ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                if (++stage == kConvStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(acc_bar);
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: lane = output position ----------------
        const int ew = warp & 3;
        const int p = p0 + ew * 32 + lane;
        const int HpWp = (prm.H + 1) * Wp;
        const int img = p / HpWp;
        const int rem = p - img * HpWp;
        const int r = rem / Wp;
        const int c = rem - r * Wp;
        const bool valid = p < prm.P && r < prm.H && c < prm.W;
        const int ldo = prm.ld_out ? prm.ld_out : prm.cout;
        const int ldr = prm.ld_res ? prm.ld_res : prm.cout;
        // destinations
        size_t o_main = 0, o_sub = 0, o_res = 0;
        bool sub_ok = false;
        if (valid) {
            if (prm.out_mode == kOutPhaseSplit) {
                const int Wh = (prm.W >> 1) + 1, HhWh = ((prm.H >> 1) + 1) * Wh;
                const int ph = ((r & 1) << 1) | (c & 1);
                o_main = static_cast<size_t>(ph) * prm.out_phase_rows + static_cast<size_t>(img) * HhWh + (r >> 1) * Wh + (c >> 1);
            } else {
                o_main = static_cast<size_t>(p);
            }
            if (prm.out_sub && !(r & 1) && !(c & 1)) {
                const int Wh = (prm.W >> 1) + 1, HhWh = ((prm.H >> 1) + 1) * Wh;
                o_sub = static_cast<size_t>(img) * HhWh + (r >> 1) * Wh + (c >> 1);
                sub_ok = true;
            }
            if (prm.res_mode == kResSame) {
                o_res = static_cast<size_t>(p);
            } else if (prm.res_mode == kResSubsample) {
                const int W2p = 2 * prm.W + 1, H2pW2p = (2 * prm.H + 1) * W2p;
                o_res = static_cast<size_t>(img) * H2pW2p + (2 * r) * W2p + 2 * c;
            } else if (prm.res_mode == kResUpsample) {
                const int Wh = (prm.W >> 1) + 1, HhWh = ((prm.H >> 1) + 1) * Wh;
                o_res = static_cast<size_t>(img) * HhWh + (r >> 1) * Wh + (c >> 1);
            }
        }
        // the residual row of this position is fetched while the tensor pipe is still busy with the main loop
        uint4 resv[BN / 8];
        if (valid && prm.res_mode != kResNone) {
            const uint4* rp = reinterpret_cast<const uint4*>(prm.res + o_res * ldr + n0);
#pragma unroll
            for (int j = 0; j < BN / 8; j += 2) ld_global_nc_256(rp + j, resv[j], resv[j + 1]);
        }
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll
        for (int cc = 0; cc < BN; cc += 16) {
            uint32_t raw[16];
            tmem_ld_32x32b_x16(taddr + cc, raw);
            tmem_ld_wait_x16(raw);
            const int n = n0 + cc;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
            if (prm.partial) {
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(prm.partial + (static_cast<size_t>(blockIdx.z) * prm.P + p) * prm.cout + n);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += s_bias[cc + j];
            if (prm.prelu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * s_prelu[cc + j];
            }
            if (prm.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (POOL) {  // every lane takes part in the shuffles
                const int pw = p0 + ew * 32;
                pool_store16(prm.pool, prm.cout, pw >> 5, n, v, valid, lane, (pw + 31) / HpWp != pw / HpWp, img == pw / HpWp);
            }
            if (!valid) continue;
            if (HEADS) {
                // columns 0-7 BboxHead (anchor*4 + k), 8-11 ClassHead (anchor*2 + class), 12-31 LandmarkHead (anchor*10 + k); bias added
                // above; 2-way softmax (retinaface_trim.py:123-127); anchor-major scatter = permute(0,2,3,1).view(B,-1,k) with the levels
                // concatenated (retinaface_trim.py:31-35,119-121)
                const size_t a0 = static_cast<size_t>(img) * prm.anchors_total + prm.level_offset + static_cast<size_t>(r * prm.W + c) * 2;
                if (cc == 0) {
                    float4* lp = reinterpret_cast<float4*>(prm.head_loc + a0 * 4);
                    lp[0] = make_float4(v[0], v[1], v[2], v[3]);
                    lp[1] = make_float4(v[4], v[5], v[6], v[7]);
                    float4 cf;
                    {
                        const float m = fmaxf(v[8], v[9]);
                        const float e0 = expf(v[8] - m), e1 = expf(v[9] - m);
                        const float inv = 1.f / (e0 + e1);
                        cf.x = e0 * inv;
                        cf.y = e1 * inv;
                    }
                    {
                        const float m = fmaxf(v[10], v[11]);
                        const float e0 = expf(v[10] - m), e1 = expf(v[11] - m);
                        const float inv = 1.f / (e0 + e1);
                        cf.z = e0 * inv;
                        cf.w = e1 * inv;
                    }
                    *reinterpret_cast<float4*>(prm.head_conf + a0 * 2) = cf;
                    if (prm.head_landm) {
#pragma unroll
                        for (int j = 12; j < 16; ++j) prm.head_landm[a0 * 10 + (j - 12)] = v[j];
                    }
                } else if (prm.head_landm) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) prm.head_landm[a0 * 10 + (cc + j - 12)] = v[j];  // a0*10 + l*10 + k with l*10 + k = column - 12
                }
                continue;
            }
            if (prm.res_mode != kResNone) {
                const uint4 r0 = resv[cc / 8], r1 = resv[cc / 8 + 1];
                const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
                const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 a = __half22float2(h0[j]), b = __half22float2(h1[j]);
                    v[2 * j] += a.x;
                    v[2 * j + 1] += a.y;
                    v[8 + 2 * j] += b.x;
                    v[8 + 2 * j + 1] += b.y;
                }
            }
            uint4 pk[2];
            __half2* hp = reinterpret_cast<__half2*>(pk);
#pragma unroll
            for (int j = 0; j < 8; ++j) hp[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            if (prm.out2 && n >= prm.out2_from) {
                st_global_256(prm.out2 + o_main * prm.ld_out2 + (n - prm.out2_from), pk[0], pk[1]);
            } else if (prm.out) {
                st_global_256(prm.out + o_main * ldo + n, pk[0], pk[1]);
            }
            if (sub_ok) {
                st_global_256(prm.out_sub + o_sub * ldo + n, pk[0], pk[1]);
            }
            if (prm.out_bn) {
                // the stored (fp16-rounded) value is what the next unit's shortcut sees; its BN input is the same value
                uint4 pb[2];
                __half2* hb = reinterpret_cast<__half2*>(pb);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 y = __half22float2(hp[j]);
                    hb[j] = __floats2half2_rn(fmaf(y.x, s_bns[cc + 2 * j], s_bnb[cc + 2 * j]), fmaf(y.y, s_bns[cc + 2 * j + 1], s_bnb[cc + 2 * j + 1]));
                }
                st_global_256(prm.out_bn + o_main * ldo + n, pb[0], pb[1]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace frb
