// FUSED conv_dw block of MobileNetV1-0.25 for >= 64 input channels (/root/reference conversion/retina/models/net.py:29-38):
// depthwise 3x3 (stride 1 or 2, pad 1) + BN + ReLU on CUDA cores, feeding the pointwise 1x1 + BN + ReLU as a tcgen05 GEMM in the same
// kernel. The depthwise result never goes to global memory: four producer warps compute a [128 positions x 64 channels] block of it
// straight into shared memory in the tensor core's operand layout (K-major rows of 128 B, 128-byte swizzle: 16-byte chunk j of row r
// lives at chunk j ^ (r & 7)), publish it with fence.proxy.async + an mbarrier, and the MMA warp multiplies it with the pointwise
// weights that a TMA warp streams in. Replaces dw3x3_kernel + conv_gemm_kernel per block: one launch instead of two, and the
// depthwise map's write + re-read disappear.
//   warp 0: TMA (pointwise weights, one [BN x 64] box per 64-channel block)     warp 1: MMA issuer (M = 128, N = BN = Cout)
//   warp 2: TMEM allocator                                                       warps 4-7: depthwise producers, then epilogue
// Activations: fp16 shared-halo flat NHWC (conv_kernels.cuh). dw_w: [9][Cin] f32, dw_b: [Cin], pw weights [Cout][Cin] fp16, pw_b: [Cout].
#pragma once
#include "conv_kernels.cuh"
#include "det_kernels.cuh"

namespace frb {

struct DwPwGemmParams {
    const __half* in;
    Geo gi, go;
    int stride;  // 1 or 2
    int cin;     // multiple of 64
    int P;       // output positions = batch * go.HpWp()
    const float* dw_w;
    const float* dw_b;
    const float* pw_b;
    __half* out;  // [P, BN]
};

template <int BN>
struct DwPwGemmCfg {
    static constexpr int kABytes = kConvBM * 128;
    static constexpr int kBBytes = BN * 128;
    static constexpr int kStages = 2;
    // [A stages][B stages][barriers 256][dw weights + bias: 10 * cin floats, cin <= 256][pw bias BN]
    static constexpr int smem_bytes(int cin) { return 1024 + kStages * (kABytes + kBBytes) + 256 + 10 * cin * 4 + BN * 4; }
};

template <int BN>
__global__ void __launch_bounds__(256) dwpw_gemm_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ DwPwGemmParams prm) {
    using Cfg = DwPwGemmCfg<BN>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + kStages * Cfg::kABytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + kStages * Cfg::kBBytes);
    uint64_t* a_empty = a_full + kStages;
    uint64_t* b_full = a_empty + kStages;
    uint64_t* b_empty = b_full + kStages;
    uint64_t* acc_bar = b_empty + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    float* s_dw = reinterpret_cast<float*>(b_ring + kStages * Cfg::kBBytes + 256);  // [9][cin]
    float* s_db = s_dw + 9 * prm.cin;                                                 // [cin]
    float* s_pb = s_db + prm.cin;                                                     // [BN]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int p0 = blockIdx.x * kConvBM;
    const int cin_blocks = prm.cin >> 6;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_b);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&a_full[s], 128);  // every producer thread arrives
            mbar_init(&a_empty[s], 1);
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<BN>(tmem_slot);
    if (warp >= 4) {
        const int t = threadIdx.x - 128;
        for (int i = t; i < 9 * prm.cin; i += 128) s_dw[i] = __ldg(prm.dw_w + i);
        for (int i = t; i < prm.cin; i += 128) s_db[i] = __ldg(prm.dw_b + i);
        for (int i = t; i < BN; i += 128) s_pb[i] = __ldg(prm.pw_b + i);
    }
    griddep_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();

    if (warp == 0) {
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int cb = 0; cb < cin_blocks; ++cb) {
                mbar_wait(&b_empty[stage], phase ^ 1);
                mbar_expect_tx(&b_full[stage], Cfg::kBBytes);
                tma_load_2d(b_ring + stage * Cfg::kBBytes, &tmap_b, &b_full[stage], cb * 64, 0, kEvictLast);
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(kConvBM, BN, 0, 0);
            uint32_t stage = 0, phase = 0;
            for (int cb = 0; cb < cin_blocks; ++cb) {
                mbar_wait(&a_full[stage], phase);
                mbar_wait(&b_full[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(a_ring + stage * Cfg::kABytes);
                const uint32_t b_addr = smem_u32(b_ring + stage * Cfg::kBBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16_ss(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (cb > 0 || k > 0) ? 1u : 0u);
                umma_commit(&a_empty[stage]);
                umma_commit(&b_empty[stage]);
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(acc_bar);
        }
    } else if (warp >= 4) {
        // ---------------- depthwise producers: thread = (16-byte channel chunk j, row group rg); rows rg, rg + 16, ... ----------------
        const int t = threadIdx.x - 128;
        const int j = t & 7, rg = t >> 3;
        const int Wpi = prm.gi.Wp(), HpWpi = prm.gi.HpWp(), Wpo = prm.go.Wp(), HpWpo = prm.go.HpWp();
        int in_off[8];      // matrix row of the centre input pixel, -1: not an output pixel
        uint32_t edge = 0;  // bit i: top row missing (ri == 0), bit 8 + i: left column missing (ci == 0)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = p0 + rg + 16 * i;
            const int img = p / HpWpo, rem = p - img * HpWpo;
            const int r = rem / Wpo, c = rem - r * Wpo;
            const bool valid = p < prm.P && r < prm.go.H && c < prm.go.W;
            const int ri = r * prm.stride, ci = c * prm.stride;
            in_off[i] = valid ? img * HpWpi + ri * Wpi + ci : -1;
            if (ri == 0) edge |= 1u << i;
            if (ci == 0) edge |= 1u << (8 + i);
        }
        uint32_t stage = 0, phase = 0;
        for (int cb = 0; cb < cin_blocks; ++cb) {
            const int ch0 = cb * 64 + j * 8;
            mbar_wait(&a_empty[stage], phase ^ 1);
            uint8_t* a_dst = a_ring + stage * Cfg::kABytes;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = rg + 16 * i;
                uint4 pk = make_uint4(0, 0, 0, 0);
                if (in_off[i] >= 0) {
                    float acc[8];
                    {
                        const float4 b0 = *reinterpret_cast<const float4*>(s_db + ch0), b1 = *reinterpret_cast<const float4*>(s_db + ch0 + 4);
                        acc[0] = b0.x, acc[1] = b0.y, acc[2] = b0.z, acc[3] = b0.w, acc[4] = b1.x, acc[5] = b1.y, acc[6] = b1.z, acc[7] = b1.w;
                    }
                    const __half* centre = prm.in + static_cast<size_t>(in_off[i]) * prm.cin + ch0;
                    uint4 v[9];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            // the row below / column right of the map are the layout's zero pads (read as stored); above / left are skipped
                            const bool ok = !(ky == 0 && ((edge >> i) & 1u)) && !(kx == 0 && ((edge >> (8 + i)) & 1u));
                            v[ky * 3 + kx] = ok ? __ldg(reinterpret_cast<const uint4*>(centre + (static_cast<ptrdiff_t>(ky - 1) * Wpi + (kx - 1)) * prm.cin))
                                                : make_uint4(0, 0, 0, 0);
                        }
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const float4 w0 = *reinterpret_cast<const float4*>(s_dw + tap * prm.cin + ch0);
                        const float4 w1 = *reinterpret_cast<const float4*>(s_dw + tap * prm.cin + ch0 + 4);
                        const __half2* h = reinterpret_cast<const __half2*>(&v[tap]);
                        const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), d = __half22float2(h[2]), e = __half22float2(h[3]);
                        acc[0] = fmaf(a.x, w0.x, acc[0]);
                        acc[1] = fmaf(a.y, w0.y, acc[1]);
                        acc[2] = fmaf(b.x, w0.z, acc[2]);
                        acc[3] = fmaf(b.y, w0.w, acc[3]);
                        acc[4] = fmaf(d.x, w1.x, acc[4]);
                        acc[5] = fmaf(d.y, w1.y, acc[5]);
                        acc[6] = fmaf(e.x, w1.z, acc[6]);
                        acc[7] = fmaf(e.y, w1.w, acc[7]);
                    }
                    __half2* hp = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                    for (int q = 0; q < 4; ++q) hp[q] = __floats2half2_rn(fmaxf(acc[2 * q], 0.f), fmaxf(acc[2 * q + 1], 0.f));
                }
                *reinterpret_cast<uint4*>(a_dst + row * 128 + ((j ^ (row & 7)) << 4)) = pk;
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
            mbar_arrive(&a_full[stage]);
            if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
            }
        }
        // ---------------- epilogue: lane = output position ----------------
        const int ew = warp & 3;
        const int p = p0 + ew * 32 + lane;
        const int img = p / HpWpo, rem = p - img * HpWpo;
        const int r = rem / Wpo, c = rem - r * Wpo;
        const bool valid = p < prm.P && r < prm.go.H && c < prm.go.W;
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll
        for (int cc = 0; cc < BN; cc += 16) {
            uint32_t raw[16];
            tmem_ld_32x32b_x16(taddr + cc, raw);
            tmem_ld_wait_x16(raw);
            if (!valid) continue;
            uint4 pk[2];
            __half2* hp = reinterpret_cast<__half2*>(pk);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                hp[q] = __floats2half2_rn(fmaxf(__uint_as_float(raw[2 * q]) + s_pb[cc + 2 * q], 0.f),
                                          fmaxf(__uint_as_float(raw[2 * q + 1]) + s_pb[cc + 2 * q + 1], 0.f));
            st_global_256(prm.out + static_cast<size_t>(p) * BN + cc, pk[0], pk[1]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<BN>(tmem_base);
}

}  // namespace frb
