// 3x3 stride-1 convolution on tcgen05, PERSISTENT, HALO-REUSING, MULTI-TILE (the body convs of ArcFace IR-(SE)50:
// /root/reference conversion/arcface/model_irse.py:48-90). Same activation layout, tensor maps, parameters and epilogue arithmetic as
// conv_gemm_kernel (conv_kernels.cuh); what changes is how the operands reach the tensor core:
//
//   * conv_gemm_kernel fetches, per 64-channel k-block, one [128 x 64] activation box AND one [BN x 64] weight box for 4 MMAs: at
//     BN = 128 that is 32 KiB of L2 -> SM traffic per 256 tensor cycles = 128 B/cycle/SM, three times what the L2 delivers
//     (~6300 B/cycle chip-wide = 42.5 B/cycle/SM): the kernel is L2-bound at a third of the tensor peak (measured 0.4-1.0 PFLOP/s).
//   * here one work unit is MT adjacent 128-position tiles x BN channels. Per 64-channel block ONE halo tile (the MT*128 positions plus
//     a row and a column of neighbours on each side) is fetched and all nine taps of all MT tiles are row-shifted UMMA descriptors
//     into it; each weight box feeds MT*4 MMAs. BN = 128, MT = 2: (16 + 5.3) KiB per 512 tensor cycles = 43 B/cycle/SM.
//     With 64 -> 64 channels the nine weight boxes (72 KiB) stay resident for the whole kernel (weight-stationary).
//   * persistent: grid = min(units, SMs); two TMEM accumulator sets are ping-ponged so the epilogue of unit i (8 warps) overlaps the
//     MMAs of unit i + 1; halo tiles are double buffered, weights stream through a ring.
//   * optional SE pooling: per 32-position group the epilogue also writes the exact (fixed-point) channel sums of the conv output
//     (warp shuffle tree), which se_gate_apply_kernel reduces per image - the separate pooling pass over the map disappears.
#pragma once
#include "conv_kernels.cuh"

namespace frb {

constexpr int kMtThreads = 128 + 8 * 32;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr int kMtMaxStages = 12;
constexpr int kMtParamBytes = 4 * 512 * 4 + 256;  // bias, prelu, bn_s, bn_b for up to 512 output channels + the tap tables

// The nine taps of a 64-channel block are organised in GROUPS that share one halo tile:
//   stride 1: one group, halo tile = matrix rows p0 - Wp - 1 ... of the input, tap (dy, dx) starts dy*Wp + dx rows in;
//   stride 2 (phase-split input, see conv_kernels.cuh): one group per phase map (1 + 2 + 2 + 4 taps), halo tile = rows
//   p0 - Wp - 1 ... of that phase map, tap (dy, dx) starts Wp + 1 - (dy == 0 ? Wp : 0) - (dx == 0 ? 1 : 0) rows in.
struct MtGroup {
    int row0;    // the halo tile starts at matrix row p0 + row0
    int ntaps;
    int tap[4 + 5];  // weight tap index (dy*3 + dx)
    int off[4 + 5];  // operand start row inside the halo tile
};
struct ConvMtExtra {
    int units;        // ceil(P / (MT*128)) * (cout / BN)
    int n_blocks;     // cout / BN
    int stages;       // weight ring depth (<= kMtMaxStages)
    int halo_chunks;  // 128-row boxes per halo tile
    int stationary;   // 1: stride 1, cin_blocks == 1 && n_blocks == 1 && stages >= 9: weights loaded once per CTA
    int ngroups;      // 1 (stride 1) or 4 (stride 2)
    // conv3x3_pair_kernel only: work items [0, full_units) are whole units; the units of a short last round are split into two
    // half-width items each (channels [0, BN/2) and [BN/2, BN) of the unit), so that the round costs half a unit's time
    int full_units, half_items;
    MtGroup grp[4];
};

// shared-memory bytes of one CTA
inline int conv_mt_smem_bytes(int BN, int halo_chunks, int stages) {
    return 1024 + 2 * halo_chunks * kConvBM * 128 + stages * BN * 128 + 512 + kMtParamBytes;
}

// Epilogue of one warp: its 32 output positions pw .. pw + 31 (lane = position = TMEM lane) x kCols channels starting at n0, accumulator
// columns at `taddr`. Waits for the accumulator (acc_full / acc_parity) after the residual rows have been requested, then bias, PReLU /
// ReLU, SE pooling partials, residual, fp16 stores (main / phase-split / subsampled / next-unit-BN destinations). Every tcgen05.ld has
// completed on return. s_*: the layer's per-channel parameters indexed by GLOBAL channel.
template <int kCols>
__device__ __forceinline__ void conv_epilogue_rows(const ConvGemmParams& prm, const float* s_bias, const float* s_prelu, const float* s_bns,
                                                   const float* s_bnb, uint32_t taddr, int pw, int n0, int lane, uint64_t* acc_full,
                                                   uint32_t acc_parity) {
    const int Wp = prm.W + 1;
    const int HpWp = (prm.H + 1) * Wp;
    const int ldo = prm.ld_out ? prm.ld_out : prm.cout;
    const int ldr = prm.ld_res ? prm.ld_res : prm.cout;
    const int p = pw + lane;
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
    const int img_w0 = pw / HpWp;
    const bool straddles = (pw + 31) / HpWp != img_w0;
    // the residual row of this position is fetched while the tensor pipe is still busy with this unit
    uint4 resv[kCols / 8];
    if (valid && prm.res_mode != kResNone) {
        const uint4* rp = reinterpret_cast<const uint4*>(prm.res + o_res * ldr + n0);
#pragma unroll
        for (int j = 0; j < kCols / 8; j += 2) ld_global_nc_256(rp + j, resv[j], resv[j + 1]);
    }
    mbar_wait(acc_full, acc_parity);
    tc_fence_after();
#pragma unroll
    for (int cc = 0; cc < kCols; cc += 16) {
        uint32_t raw[16];
        tmem_ld_32x32b_x16(taddr + cc, raw);
        tmem_ld_wait_x16(raw);
        const int n = n0 + cc;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) + s_bias[n + j];
        if (prm.prelu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * s_prelu[n + j];
        }
        if (prm.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (prm.pool)  // warp-uniform: every lane takes part in the shuffles
            pool_store16(prm.pool, prm.cout, pw >> 5, n, v, valid, lane, straddles, img == img_w0);
        if (!valid) continue;
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
            st_global_256(prm.out + o_main * ldo + n, pk[0], pk[1]);
        }
        if (sub_ok) {
            st_global_256(prm.out_sub + o_sub * ldo + n, pk[0], pk[1]);
        }
        if (prm.out_bn) {
            // the stored (fp16-rounded) value is what the next unit's shortcut sees; its BN input is the same value
            uint4 pb[2];
            __half2* hb2 = reinterpret_cast<__half2*>(pb);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 y = __half22float2(hp[j]);
                hb2[j] = __floats2half2_rn(fmaf(y.x, s_bns[n + 2 * j], s_bnb[n + 2 * j]), fmaf(y.y, s_bns[n + 2 * j + 1], s_bnb[n + 2 * j + 1]));
            }
            st_global_256(prm.out_bn + o_main * ldo + n, pb[0], pb[1]);
        }
    }
}

template <int BN, int MT>
__global__ void __launch_bounds__(kMtThreads, 1)
conv3x3_mt_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ ConvGemmParams prm, const __grid_constant__ ConvMtExtra ex) {
    griddep_launch_dependents();
    constexpr int kABytes = kConvBM * 128, kBBytes = BN * 128;
    constexpr uint32_t kTmemCols = 2 * MT * BN;  // two accumulator sets
    static_assert(kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM allocation must be a power of two <= 512");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    // [2 halo tiles][weight ring][barriers][epilogue parameters of all cout channels]
    const int halo_bytes = ex.halo_chunks * kABytes;
    uint8_t* ring = smem + 2 * halo_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + ex.stages * kBBytes);
    uint64_t* empty_bar = full_bar + kMtMaxStages;
    uint64_t* hfull_bar = empty_bar + kMtMaxStages;  // [2] halo tile landed
    uint64_t* hempty_bar = hfull_bar + 2;             // [2] halo tile consumed by all its MMAs
    uint64_t* tfull_bar = hempty_bar + 2;             // [2] accumulator set complete
    uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator set read out by the eight epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_bias = reinterpret_cast<float*>(ring + ex.stages * kBBytes + 512);
    float* s_prelu = s_bias + 512;
    float* s_bns = s_prelu + 512;
    float* s_bnb = s_bns + 512;
    // tap tables of the MMA issuer (the issue loop must stay far below the 256-512 tensor cycles of a tap: no indexed constant loads)
    uint32_t* s_off8 = reinterpret_cast<uint32_t*>(s_bnb + 512);  // [4][9] operand start inside the halo tile, in 16-byte units
    uint32_t* s_ntaps = s_off8 + 36;                               // [4]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int Wp = prm.W + 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kMtMaxStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&hfull_bar[b], 1);
            mbar_init(&hempty_bar[b], 1);
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 8);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
    if (warp == 3) {
        for (int i = lane; i < 36; i += 32) s_off8[i] = static_cast<uint32_t>(ex.grp[i / 9].off[i % 9]) * 8u;
        if (lane < 4) s_ntaps[lane] = static_cast<uint32_t>(ex.grp[lane].ntaps);
    }
    if (warp >= 4) {
        for (int i = threadIdx.x - 128; i < prm.cout; i += kMtThreads - 128) {
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
    griddep_wait();  // everything above touched only this CTA's state and static parameters

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (elect_one()) {
            uint32_t stage = 0, phase = 0, hb = 0, hph = 0;
            bool first = true;
            for (int u = blockIdx.x; u < ex.units; u += gridDim.x) {
                const int pt = u / ex.n_blocks, nb = u - pt * ex.n_blocks;
                const int p0 = pt * (MT * kConvBM), n0 = nb * BN;
                for (int cb = 0; cb < prm.cin_blocks; ++cb) {
                    for (int g = 0; g < ex.ngroups; ++g) {
                        const MtGroup& G = ex.grp[g];
                        mbar_wait(&hempty_bar[hb], hph ^ 1);
                        mbar_expect_tx(&hfull_bar[hb], halo_bytes);
                        for (int ch = 0; ch < ex.halo_chunks; ++ch)
                            tma_load_2d(smem + hb * halo_bytes + ch * kABytes, &tmap_a, &hfull_bar[hb], cb * 64, p0 + G.row0 + ch * kConvBM, kEvictNormal);
                        if (++hb == 2) {
                            hb = 0;
                            hph ^= 1;
                        }
                        if (ex.stationary && !first) continue;
                        for (int t = 0; t < G.ntaps; ++t) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            mbar_expect_tx(&full_bar[stage], kBBytes);
                            tma_load_2d(ring + stage * kBBytes, &tmap_b, &full_bar[stage], (G.tap[t] * prm.cin_blocks + cb) * 64, n0, kEvictLast);
                            if (++stage == static_cast<uint32_t>(ex.stages)) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
                first = false;
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(kConvBM, BN, 0, 0);
            // shared-memory matrix descriptors (umma_desc_sw128) as (hi, lo) words: hi is constant, lo = flags | (address >> 4); stepping
            // to another tap / tile / k-slice / ring stage is an add on the low word
            constexpr uint32_t kDescHi = 0x40004040u;  // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
            constexpr uint32_t kDescLo = 0x10000u;     // LBO = 1
            auto desc = [](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
            const uint32_t a_lo0 = kDescLo | ((smem_u32(smem) & 0x3FFFFu) >> 4);
            const uint32_t a_lo_step = static_cast<uint32_t>(halo_bytes) >> 4;
            const uint32_t b_lo0 = kDescLo | ((smem_u32(ring) & 0x3FFFFu) >> 4);
            uint32_t stage = 0, phase = 0, hb = 0, hph = 0;
            bool first = true;
            int i = 0;
            for (int u = blockIdx.x; u < ex.units; u += gridDim.x, ++i) {
                const uint32_t buf = i & 1;
                mbar_wait(&tempty_bar[buf], ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (MT * BN);
                uint32_t acc = 0;  // 0 for the very first MMA into each accumulator of this unit
                for (int cb = 0; cb < prm.cin_blocks; ++cb) {
                    for (int g = 0; g < ex.ngroups; ++g) {
                        const uint32_t ntaps = s_ntaps[g];
                        const uint32_t* off8 = s_off8 + g * 9;
                        mbar_wait(&hfull_bar[hb], hph);
                        tc_fence_after();
                        const uint32_t a_lo_h = a_lo0 + hb * a_lo_step;
#pragma unroll 1
                        for (uint32_t t = 0; t < ntaps; ++t) {
                            if (!ex.stationary || first) {
                                mbar_wait(&full_bar[stage], phase);
                                tc_fence_after();
                            }
                            // the tap's operand for tile m = 128 consecutive rows of the halo tile starting m*128 + off rows in (the swizzle acts
                            // on absolute shared-memory address bits, so a 128-byte-row offset needs no descriptor base offset: measured)
                            const uint32_t a_lo = a_lo_h + off8[t];
                            const uint32_t b_lo = b_lo0 + stage * (kBBytes >> 4);
#pragma unroll
                            for (int m = 0; m < MT; ++m) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_f16_ss(d_tmem + m * BN, desc(a_lo + m * (kABytes >> 4) + k * 2), desc(b_lo + k * 2), idesc, (acc | k) ? 1u : 0u);
                            }
                            acc = 1;
                            if (!ex.stationary) umma_commit(&empty_bar[stage]);
                            if (++stage == static_cast<uint32_t>(ex.stages)) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                        umma_commit(&hempty_bar[hb]);
                        if (++hb == 2) {
                            hb = 0;
                            hph ^= 1;
                        }
                    }
                }
                umma_commit(&tfull_bar[buf]);
                first = false;
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: lane = output position; warps 4-7 / 8-11 split the unit ----------------
        const int ew = warp & 3;          // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;  // MT == 2: tile of the unit; MT == 1: column half
        const int m = MT == 2 ? half : 0;
        constexpr int kCols = MT == 2 ? BN : BN / 2;  // columns this warp handles
        const int c0 = MT == 2 ? 0 : half * (BN / 2);
        int i = 0;
        for (int u = blockIdx.x; u < ex.units; u += gridDim.x, ++i) {
            const uint32_t buf = i & 1;
            const int pt = u / ex.n_blocks, nb = u - pt * ex.n_blocks;
            const int n0 = nb * BN + c0;
            const int pw = pt * (MT * kConvBM) + m * kConvBM + ew * 32;  // first position of this warp
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * (MT * BN) + m * BN + c0;
            conv_epilogue_rows<kCols>(prm, s_bias, s_prelu, s_bns, s_bnb, taddr, pw, n0, lane, &tfull_bar[buf], (i >> 1) & 1);
            // every tcgen05.ld of this warp has completed (tmem_ld_wait_x16 above): hand the accumulator set back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// The same convolution on a CTA PAIR (cta_group::2, cluster of two CTAs on the two SMs of a TPC). conv3x3_mt_kernel is bound by
// shared-memory bandwidth once the operands come from L2 fast enough: a single-CTA MMA (M = 128, N = 128, K = 16) reads 4 KiB of A
// and 4 KiB of B per 64 tensor cycles - the whole 128 B/cycle of an SM's shared memory - while TMA writes the next operands into the
// same memory (ncu: tensor pipe 60-64 % of the active cycles on the 128 / 256-channel layers). In a pair the MMA is M = 256 (128
// positions in each CTA) x N = BN, each CTA holds its own halo tile and only its HALF of every weight tile (BN / 2 rows), and the
// hardware feeds both tensor cores from both halves: per SM 4 KiB of A + 4 KiB of B per 128 tensor cycles at BN = 256.
//   work unit  = 256 positions (CTA rank r: positions p0 + 128 r ...) x BN channels; two accumulator sets of BN columns in TMEM
//   producers  = warp 0 of EACH CTA (its halo tile + its half of the weight tile; completion bytes go to the LEADER's barriers)
//   MMA issuer = warp 1 of the leader CTA only (tcgen05.mma.cta_group::2); commits are multicast to both CTAs
//   epilogue   = warps 4-11 of each CTA for its own 128 positions (same code as conv3x3_mt_kernel)
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMtThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_b_half, const __grid_constant__ ConvGemmParams prm,
                    const __grid_constant__ ConvMtExtra ex) {
    griddep_launch_dependents();
    constexpr int kABytes = kConvBM * 128, kBHalf = (BN / 2) * 128;
    constexpr uint32_t kTmemCols = 2 * BN;  // two accumulator sets
    static_assert(kTmemCols == 256 || kTmemCols == 512, "TMEM allocation must be a power of two <= 512");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    // [2 halo tiles][weight ring (this CTA's half tiles)][barriers][epilogue parameters][tap tables]: identical offsets in both CTAs
    const int halo_bytes = ex.halo_chunks * kABytes;
    uint8_t* ring = smem + 2 * halo_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + ex.stages * kBHalf);  // leader: both halves of a weight tile landed
    uint64_t* empty_bar = full_bar + kMtMaxStages;                                 // each CTA: its ring slot was consumed
    uint64_t* hfull_bar = empty_bar + kMtMaxStages;                                // leader: both CTAs' halo tiles landed
    uint64_t* hempty_bar = hfull_bar + 2;                                          // each CTA
    uint64_t* tfull_bar = hempty_bar + 2;                                          // each CTA: accumulator set complete
    uint64_t* tempty_bar = tfull_bar + 2;                                          // leader: set read out by the 16 epilogue warps of the pair
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_bias = reinterpret_cast<float*>(ring + ex.stages * kBHalf + 512);
    float* s_prelu = s_bias + 512;
    float* s_bns = s_prelu + 512;
    float* s_bnb = s_bns + 512;
    uint32_t* s_off8 = reinterpret_cast<uint32_t*>(s_bnb + 512);
    uint32_t* s_ntaps = s_off8 + 36;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int Wp = prm.W + 1;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int items = ex.full_units + ex.half_items;
    // work item w -> (unit, first channel, width): whole units first, then the half-width items of the short last round
    auto decode = [&](int w, int& pt, int& n_base, bool& half) {
        half = w >= ex.full_units;
        const int u = half ? ex.full_units + ((w - ex.full_units) >> 1) : w;
        pt = u / ex.n_blocks;
        n_base = (u - pt * ex.n_blocks) * BN + (half ? ((w - ex.full_units) & 1) * (BN / 2) : 0);
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_b_half);
    }
    if (warp == 1 && lane == 0) {
        for (int s2 = 0; s2 < kMtMaxStages; ++s2) {
            mbar_init(&full_bar[s2], 1);
            mbar_init(&empty_bar[s2], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&hfull_bar[b], 1);
            mbar_init(&hempty_bar[b], 1);
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 16);  // one arrive per epilogue warp of both CTAs
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_pair<kTmemCols>(tmem_slot);
    if (warp == 3) {
        for (int i = lane; i < 36; i += 32) s_off8[i] = static_cast<uint32_t>(ex.grp[i / 9].off[i % 9]) * 8u;
        if (lane < 4) s_ntaps[lane] = static_cast<uint32_t>(ex.grp[lane].ntaps);
    }
    if (warp >= 4) {
        for (int i = threadIdx.x - 128; i < prm.cout; i += kMtThreads - 128) {
            s_bias[i] = prm.bias ? __ldg(prm.bias + i) : 0.f;
            s_prelu[i] = prm.prelu ? __ldg(prm.prelu + i) : 1.f;
            s_bns[i] = prm.out_bn ? __ldg(prm.bn_s + i) : 1.f;
            s_bnb[i] = prm.out_bn ? __ldg(prm.bn_b + i) : 0.f;
        }
    }
    // barriers and TMEM are ready in both CTAs before anything is signalled across them
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs) ----------------
        if (elect_one()) {
            const uint32_t leader_full0 = mapa_u32(smem_u32(&full_bar[0]), 0);
            const uint32_t leader_hfull0 = mapa_u32(smem_u32(&hfull_bar[0]), 0);
            uint32_t stage = 0, phase = 0, hb = 0, hph = 0;
            for (int w = pair; w < items; w += num_pairs) {
                int pt, n_base;
                bool half;
                decode(w, pt, n_base, half);
                const int p0 = pt * (2 * kConvBM) + static_cast<int>(rank) * kConvBM;        // this CTA's 128 positions
                const int n0 = n_base + static_cast<int>(rank) * (half ? BN / 4 : BN / 2);  // this CTA's half of the item's weight rows
                const CUtensorMap* tb = half ? &tmap_b_half : &tmap_b;
                const uint32_t b_bytes = half ? kBHalf / 2 : kBHalf;
                for (int cb = 0; cb < prm.cin_blocks; ++cb) {
                    for (int g = 0; g < ex.ngroups; ++g) {
                        const MtGroup& G = ex.grp[g];
                        mbar_wait(&hempty_bar[hb], hph ^ 1);
                        if (rank == 0) mbar_expect_tx(&hfull_bar[hb], 2 * halo_bytes);
                        for (int ch = 0; ch < ex.halo_chunks; ++ch)
                            tma_load_2d_pair(smem + hb * halo_bytes + ch * kABytes, &tmap_a, leader_hfull0 + hb * 8, cb * 64,
                                             p0 + G.row0 + ch * kConvBM, kEvictNormal);
                        if (++hb == 2) {
                            hb = 0;
                            hph ^= 1;
                        }
                        for (int t = 0; t < G.ntaps; ++t) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * b_bytes);
                            tma_load_2d_pair(ring + stage * kBHalf, tb, leader_full0 + stage * 8, (G.tap[t] * prm.cin_blocks + cb) * 64, n0,
                                             kEvictLast);
                            if (++stage == static_cast<uint32_t>(ex.stages)) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA, one thread) ----------------
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc_full = umma_idesc(2 * kConvBM, BN, 0, 0), idesc_half = umma_idesc(2 * kConvBM, BN / 2, 0, 0);
            constexpr uint32_t kDescHi = 0x40004040u;  // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
            constexpr uint32_t kDescLo = 0x10000u;     // LBO = 1
            auto desc = [](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
            const uint32_t a_lo0 = kDescLo | ((smem_u32(smem) & 0x3FFFFu) >> 4);
            const uint32_t a_lo_step = static_cast<uint32_t>(halo_bytes) >> 4;
            const uint32_t b_lo0 = kDescLo | ((smem_u32(ring) & 0x3FFFFu) >> 4);
            uint32_t stage = 0, phase = 0, hb = 0, hph = 0;
            int i = 0;
            for (int w = pair; w < items; w += num_pairs, ++i) {
                const uint32_t buf = i & 1;
                const uint32_t idesc = w >= ex.full_units ? idesc_half : idesc_full;
                mbar_wait(&tempty_bar[buf], ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                uint32_t acc = 0;
                for (int cb = 0; cb < prm.cin_blocks; ++cb) {
                    for (int g = 0; g < ex.ngroups; ++g) {
                        const uint32_t ntaps = s_ntaps[g];
                        const uint32_t* off8 = s_off8 + g * 9;
                        mbar_wait(&hfull_bar[hb], hph);
                        tc_fence_after();
                        const uint32_t a_lo_h = a_lo0 + hb * a_lo_step;
#pragma unroll 1
                        for (uint32_t t = 0; t < ntaps; ++t) {
                            mbar_wait(&full_bar[stage], phase);
                            tc_fence_after();
                            const uint32_t a_lo = a_lo_h + off8[t];
                            const uint32_t b_lo = b_lo0 + stage * (kBHalf >> 4);
#pragma unroll
                            for (int k2 = 0; k2 < 4; ++k2)
                                umma_f16_ss_pair(d_tmem, desc(a_lo + k2 * 2), desc(b_lo + k2 * 2), idesc, (acc | k2) ? 1u : 0u);
                            acc = 1;
                            umma_commit_pair(&empty_bar[stage], 0x3);  // frees the slot in both CTAs
                            if (++stage == static_cast<uint32_t>(ex.stages)) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                        umma_commit_pair(&hempty_bar[hb], 0x3);
                        if (++hb == 2) {
                            hb = 0;
                            hph ^= 1;
                        }
                    }
                }
                umma_commit_pair(&tfull_bar[buf], 0x3);
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue (both CTAs): lane = output position; warps 4-7 / 8-11 take the two column halves ----------------
        const int ew = warp & 3;
        const int colhalf = (warp - 4) >> 2;
        const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);
        int i = 0;
        for (int w = pair; w < items; w += num_pairs, ++i) {
            const uint32_t buf = i & 1;
            int pt, n_base;
            bool half;
            decode(w, pt, n_base, half);
            const int pw = pt * (2 * kConvBM) + static_cast<int>(rank) * kConvBM + ew * 32;
            const uint32_t tlane = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * BN;
            if (!half) {
                const int c0 = colhalf * (BN / 2);
                conv_epilogue_rows<BN / 2>(prm, s_bias, s_prelu, s_bns, s_bnb, tlane + c0, pw, n_base + c0, lane, &tfull_bar[buf], (i >> 1) & 1);
            } else {
                const int c0 = colhalf * (BN / 4);
                conv_epilogue_rows<BN / 4>(prm, s_bias, s_prelu, s_bns, s_bnb, tlane + c0, pw, n_base + c0, lane, &tfull_bar[buf], (i >> 1) & 1);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader0 + buf * 8);
        }
    }
    tc_fence_before();
    cluster_sync_all();  // no CTA of the pair exits (or frees TMEM) while the other may still signal it or read its shared memory
    if (warp == 2) tmem_dealloc_pair<kTmemCols>(tmem_base);
}

}  // namespace frb
