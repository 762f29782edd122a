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
    // conv3x3_halo_kernel only
    int halo_chunks;       // halo tile = halo_chunks x 128 matrix rows (>= 128 + 2 Wp + 2)
    int halo_bufs;         // 1 or 2 halo buffers
    int halo_base_offset;  // 1: put the swizzle phase of the operand's first row into the descriptor's base-offset field
};

template <int BN>
struct ConvCfg {
    static constexpr int kConvStages = ConvStages<BN>::value;
    static constexpr int kABytes = kConvBM * 128;
    static constexpr int kBBytes = BN * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kParamBytes = 4 * BN * 4;  // bias, prelu, bn_s, bn_b of this CTA's channel slice (fp32)
    static constexpr int kSmemBytes = 1024 + kConvStages * kStageBytes + 256 + kParamBytes;
    static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

template <int BN>
__global__ void __launch_bounds__(kConvThreads)
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
            for (int j = 0; j < BN / 8; ++j) resv[j] = __ldg(rp + j);
        }
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll
        for (int cc = 0; cc < BN; cc += 16) {
            uint32_t raw[16];
            tmem_ld_32x32b_x16(taddr + cc, raw);
            tmem_ld_wait_x16(raw);
            if (!valid) continue;
            const int n = n0 + cc;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
            if (prm.partial) {
                float4* dst = reinterpret_cast<float4*>(prm.partial + (static_cast<size_t>(blockIdx.z) * prm.P + p) * prm.cout + n);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
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
            if (prm.out) {
                uint4* dst = reinterpret_cast<uint4*>(prm.out + o_main * ldo + n);
                dst[0] = pk[0];
                dst[1] = pk[1];
            }
            if (sub_ok) {
                uint4* dst = reinterpret_cast<uint4*>(prm.out_sub + o_sub * ldo + n);
                dst[0] = pk[0];
                dst[1] = pk[1];
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
                uint4* dst = reinterpret_cast<uint4*>(prm.out_bn + o_main * ldo + n);
                dst[0] = pb[0];
                dst[1] = pb[1];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3 stride-1 conv with HALO REUSE: same tile, epilogue and parameters as conv_gemm_kernel, but the activation operand is
// fetched once per 64-channel block as a halo tile and the nine taps are nine row-shifted views of it in shared memory, instead of
// nine TMA boxes. conv_gemm_kernel is L2-bandwidth bound on these layers (its A tile is re-read per tap: 32 KiB per k-block at
// BN = 128); this cuts the per-channel-block traffic from 9 x (16 + BN/8) KiB to (halo 20-48) + 9 x BN/8 KiB.
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kConvThreads)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ ConvGemmParams prm) {
    using Cfg = ConvCfg<BN>;
    constexpr int kConvStages = Cfg::kConvStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    // [halo buffers: halo_bufs x halo_chunks x 16 KiB][weight ring: kConvStages x BN x 128 B][barriers][epilogue params]
    const int halo_bytes = prm.halo_chunks * Cfg::kABytes;
    uint8_t* ring = smem + prm.halo_bufs * halo_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kConvStages * Cfg::kBBytes);
    uint64_t* empty_bar = full_bar + kConvStages;
    uint64_t* acc_bar = empty_bar + kConvStages;
    uint64_t* hfull_bar = acc_bar + 1;   // [2] halo tile landed
    uint64_t* hempty_bar = hfull_bar + 2;  // [2] halo tile consumed by all nine taps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hempty_bar + 2);
    float* s_bias = reinterpret_cast<float*>(ring + kConvStages * Cfg::kBBytes + 256);
    float* s_prelu = s_bias + BN;
    float* s_bns = s_prelu + BN;
    float* s_bnb = s_bns + BN;

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
        for (int b = 0; b < 2; ++b) {
            mbar_init(&hfull_bar[b], 1);
            mbar_init(&hempty_bar[b], 1);
        }
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

    if (warp == 0) {
        // TMA producer: per 64-channel block ONE halo tile (the 128 positions plus one row and one column of neighbours on each side,
        // 128 + 2 Wp + 2 matrix rows) serves all nine taps; only the weight tiles stream per tap
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int cb = 0; cb < prm.cin_blocks; ++cb) {
                const int hb = prm.halo_bufs == 2 ? (cb & 1) : 0;
                const uint32_t hphase = prm.halo_bufs == 2 ? ((cb >> 1) & 1) : (cb & 1);
                mbar_wait(&hempty_bar[hb], hphase ^ 1);
                mbar_expect_tx(&hfull_bar[hb], halo_bytes);
                for (int ch = 0; ch < prm.halo_chunks; ++ch)
                    tma_load_2d(smem + hb * halo_bytes + ch * Cfg::kABytes, &tmap_a, &hfull_bar[hb], cb * 64, p0 - Wp - 1 + ch * kConvBM,
                                kEvictNormal);
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], Cfg::kBBytes);
                    tma_load_2d(ring + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], (tap * prm.cin_blocks + cb) * 64, n0, kEvictLast);
                    if (++stage == kConvStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(kConvBM, BN, 0, 0);
            uint32_t stage = 0, phase = 0;
            for (int cb = 0; cb < prm.cin_blocks; ++cb) {
                const int hb = prm.halo_bufs == 2 ? (cb & 1) : 0;
                const uint32_t hphase = prm.halo_bufs == 2 ? ((cb >> 1) & 1) : (cb & 1);
                mbar_wait(&hfull_bar[hb], hphase);
                tc_fence_after();
                const uint32_t halo_addr = smem_u32(smem + hb * halo_bytes);
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3, dx = tap - dy * 3;
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    // the tap's operand = 128 consecutive rows of the halo tile starting dy*Wp + dx rows in (not 1024-byte aligned:
                    // the descriptor's base-offset field carries the swizzle phase of the first row)
                    const uint32_t a_addr = halo_addr + static_cast<uint32_t>(dy * Wp + dx) * 128u;
                    const uint32_t b_addr = smem_u32(ring + stage * Cfg::kBBytes);
                    const uint64_t boff = prm.halo_base_offset ? (static_cast<uint64_t>((a_addr >> 7) & 7u) << 49) : 0ull;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_base, umma_desc_sw128(a_addr + k * 32) | boff, umma_desc_sw128(b_addr + k * 32), idesc,
                                    (cb > 0 || tap > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (++stage == kConvStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&hempty_bar[hb]);
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
            for (int j = 0; j < BN / 8; ++j) resv[j] = __ldg(rp + j);
        }
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll
        for (int cc = 0; cc < BN; cc += 16) {
            uint32_t raw[16];
            tmem_ld_32x32b_x16(taddr + cc, raw);
            tmem_ld_wait_x16(raw);
            if (!valid) continue;
            const int n = n0 + cc;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
            if (prm.partial) {
                float4* dst = reinterpret_cast<float4*>(prm.partial + (static_cast<size_t>(blockIdx.z) * prm.P + p) * prm.cout + n);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
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
            if (prm.out) {
                uint4* dst = reinterpret_cast<uint4*>(prm.out + o_main * ldo + n);
                dst[0] = pk[0];
                dst[1] = pk[1];
            }
            if (sub_ok) {
                uint4* dst = reinterpret_cast<uint4*>(prm.out_sub + o_sub * ldo + n);
                dst[0] = pk[0];
                dst[1] = pk[1];
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
                uint4* dst = reinterpret_cast<uint4*>(prm.out_bn + o_main * ldo + n);
                dst[0] = pb[0];
                dst[1] = pb[1];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3 stride-1 conv, 64 -> 64 channels, WEIGHT-STATIONARY and PERSISTENT (experimental: FR_HALO=3).
// The whole weight matrix of such a layer is 9 taps x 64 x 64 fp16 = 72 KiB: every CTA loads it ONCE and then walks position tiles
// (tile = blockIdx.x, + gridDim.x, ...), fetching per tile only the halo tile of conv3x3_halo_kernel (double buffered) - 32-48 KiB of
// L2 traffic per [128 x 64] tile instead of 216 KiB (conv_gemm_kernel) or 104-120 KiB (conv3x3_halo_kernel). Two 64-column TMEM
// accumulators are ping-ponged: the epilogue of tile i (all 64 columns read to registers, accumulator released at once) overlaps
// the MMAs of tile i + 1. Same parameters, tensor maps and epilogue arithmetic as conv_gemm_kernel<64>.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWsBN = 64;
constexpr int kWsWeightBytes = 9 * kWsBN * 128;  // 72 KiB
// 8 epilogue warps (TMEM lane quarter x column half): with one persistent CTA per SM, four were not enough to write a tile's 32 KiB of
// outputs inside the 1152 MMA cycles of the next tile (measured: 8.2 vs 7.2 ms for the whole forward at batch 256)
constexpr int kWsThreads = 128 + 8 * 32;
static __global__ void __launch_bounds__(kWsThreads)  // static: this header is included by two translation units
conv3x3_ws_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ ConvGemmParams prm) {
    constexpr int BN = kWsBN;
    constexpr int kABytes = kConvBM * 128, kBBytes = BN * 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    // [halo buffers: NB x halo_chunks x 16 KiB][weights: 9 x 8 KiB][barriers][epilogue params]; NB = prm.halo_bufs (2..4) halo tiles
    // in flight: with 2 the kernel is bound by the latency of ONE 32-48 KiB halo load per tile (measured, see embedder.cu)
    const int halo_bytes = prm.halo_chunks * kABytes;
    const int NB = prm.halo_bufs;
    uint8_t* wsm = smem + NB * halo_bytes;
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(wsm + kWsWeightBytes);
    uint64_t* hfull_bar = w_bar + 1;       // [4] halo tile landed
    uint64_t* hempty_bar = hfull_bar + 4;  // [4] halo tile consumed by all nine taps
    uint64_t* tfull_bar = hempty_bar + 4;  // [2] accumulator complete
    uint64_t* tempty_bar = tfull_bar + 2;  // [2] accumulator read out by the four epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_bias = reinterpret_cast<float*>(wsm + kWsWeightBytes + 256);
    float* s_prelu = s_bias + BN;
    float* s_bns = s_prelu + BN;
    float* s_bnb = s_bns + BN;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int Wp = prm.W + 1;
    const int tiles = (prm.P + kConvBM - 1) / kConvBM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(w_bar, 1);
        for (int b = 0; b < 4; ++b) {
            mbar_init(&hfull_bar[b], 1);
            mbar_init(&hempty_bar[b], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 8);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<2 * BN>(tmem_slot);
    if (warp >= 4) {
        for (int i = threadIdx.x - 128; i < BN; i += kWsThreads - 128) {
            s_bias[i] = prm.bias ? __ldg(prm.bias + i) : 0.f;
            s_prelu[i] = prm.prelu ? __ldg(prm.prelu + i) : 1.f;
            s_bns[i] = prm.out_bn ? __ldg(prm.bn_s + i) : 1.f;
            s_bnb[i] = prm.out_bn ? __ldg(prm.bn_b + i) : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(w_bar, kWsWeightBytes);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(wsm + tap * kBBytes, &tmap_b, w_bar, tap * 64, 0, kEvictLast);
            int hb = 0;
            uint32_t hph = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(&hempty_bar[hb], hph ^ 1);
                mbar_expect_tx(&hfull_bar[hb], halo_bytes);
                for (int ch = 0; ch < prm.halo_chunks; ++ch)
                    tma_load_2d(smem + hb * halo_bytes + ch * kABytes, &tmap_a, &hfull_bar[hb], 0, t * kConvBM - Wp - 1 + ch * kConvBM,
                                kEvictNormal);
                if (++hb == NB) {
                    hb = 0;
                    hph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(kConvBM, BN, 0, 0);
            mbar_wait(w_bar, 0);
            tc_fence_after();
            const uint32_t w_addr = smem_u32(wsm);
            int i = 0, hb = 0;
            uint32_t hph = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
                const int buf = i & 1;  // accumulator
                const uint32_t ph = (i >> 1) & 1;
                mbar_wait(&tempty_bar[buf], ph ^ 1);
                tc_fence_after();
                mbar_wait(&hfull_bar[hb], hph);
                tc_fence_after();
                const uint32_t halo_addr = smem_u32(smem + hb * halo_bytes);
                const uint32_t d_tmem = tmem_base + buf * BN;
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3, dx = tap - dy * 3;
                    const uint32_t a_addr = halo_addr + static_cast<uint32_t>(dy * Wp + dx) * 128u;  // row-shifted view, base offset 0
                    const uint32_t b_addr = w_addr + tap * kBBytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (tap > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&hempty_bar[hb]);
                umma_commit(&tfull_bar[buf]);
                if (++hb == NB) {
                    hb = 0;
                    hph ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: lane = output position, warp = (TMEM lane quarter, column half) ----------------
        const int ew = warp & 3;
        const int c0 = ((warp - 4) >> 2) * (BN / 2);  // first of this warp's 32 output channels
        const int HpWp = (prm.H + 1) * Wp;
        const int ldo = prm.ld_out ? prm.ld_out : prm.cout;
        const int ldr = prm.ld_res ? prm.ld_res : prm.cout;
        int i = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
            const int buf = i & 1;
            const int p = t * kConvBM + ew * 32 + lane;
            const int img = p / HpWp;
            const int rem = p - img * HpWp;
            const int r = rem / Wp;
            const int c = rem - r * Wp;
            const bool valid = p < prm.P && r < prm.H && c < prm.W;
            size_t o_main = 0, o_sub = 0, o_res = 0;
            bool sub_ok = false;
            if (valid) {
                if (prm.out_mode == kOutPhaseSplit) {
                    const int Wh = (prm.W >> 1) + 1, HhWh = ((prm.H >> 1) + 1) * Wh;
                    const int phs = ((r & 1) << 1) | (c & 1);
                    o_main = static_cast<size_t>(phs) * prm.out_phase_rows + static_cast<size_t>(img) * HhWh + (r >> 1) * Wh + (c >> 1);
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
            uint4 resv[BN / 16];
            if (valid && prm.res_mode != kResNone) {
                const uint4* rp = reinterpret_cast<const uint4*>(prm.res + o_res * ldr + c0);
#pragma unroll
                for (int j = 0; j < BN / 16; ++j) resv[j] = __ldg(rp + j);
            }
            mbar_wait(&tfull_bar[buf], (i >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * BN + c0;
            uint32_t raw0[16], raw1[16];
            tmem_ld_32x32b_x16(taddr, raw0);
            tmem_ld_32x32b_x16(taddr + 16, raw1);
            tmem_ld_wait_x16(raw0);
            tmem_ld_wait_x16(raw1);
            // the accumulator is in registers: hand it back to the MMA warp before the arithmetic and the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
            if (!valid) continue;
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
                const uint32_t(&raw)[16] = q4 == 0 ? raw0 : raw1;
                const int cc = c0 + q4 * 16;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) + s_bias[cc + j];
                if (prm.prelu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * s_prelu[cc + j];
                }
                if (prm.relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (prm.res_mode != kResNone) {
                    const uint4 r0 = resv[q4 * 2], r1 = resv[q4 * 2 + 1];
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
                if (prm.out) {
                    uint4* dst = reinterpret_cast<uint4*>(prm.out + o_main * ldo + cc);
                    dst[0] = pk[0];
                    dst[1] = pk[1];
                }
                if (sub_ok) {
                    uint4* dst = reinterpret_cast<uint4*>(prm.out_sub + o_sub * ldo + cc);
                    dst[0] = pk[0];
                    dst[1] = pk[1];
                }
                if (prm.out_bn) {
                    uint4 pb[2];
                    __half2* hb = reinterpret_cast<__half2*>(pb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 y = __half22float2(hp[j]);
                        hb[j] = __floats2half2_rn(fmaf(y.x, s_bns[cc + 2 * j], s_bnb[cc + 2 * j]), fmaf(y.y, s_bns[cc + 2 * j + 1], s_bnb[cc + 2 * j + 1]));
                    }
                    uint4* dst = reinterpret_cast<uint4*>(prm.out_bn + o_main * ldo + cc);
                    dst[0] = pb[0];
                    dst[1] = pb[1];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<2 * BN>(tmem_base);
}

}  // namespace frb
