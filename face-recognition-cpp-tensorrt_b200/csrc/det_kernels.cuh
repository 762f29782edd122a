// CUDA-core kernels of the RetinaFace mobile0.25 detector (network spec: /root/reference conversion/retina/models/net.py,
// retinaface_trim.py, retinaface.py; pre/post-processing: src/retinaface.cpp:106-136,154-271). The 1x1 / 3x3 convolutions
// with >= 64 input channels run on the tensor cores (conv_kernels.cuh); the layers here are the HBM-bound rest:
// the 3-channel stem, the 13 depthwise convs, the four small pointwise convs, the 16-channel SSH convs, and the decode + NMS.
// Activations: fp16, shared-halo flat NHWC (see conv_kernels.cuh); accumulation fp32; BatchNorm folded by the packer.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/fr_b200.h"
#include "ptx_sm100.cuh"

namespace frb {

struct Geo {  // one feature-map geometry
    int H, W;
    __host__ __device__ int Wp() const { return W + 1; }
    __host__ __device__ int HpWp() const { return (H + 1) * (W + 1); }
};

// ---- RetinaFace::preprocess, resize + paste (src/retinaface.cpp:111-126): the frame is resized to (rw x rh) with OpenCV's u8
//      INTER_LINEAR arithmetic (11-bit fixed-point coefficients, ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2) and pasted at
//      (ox, oy) on a canvas filled with 128. One thread per canvas pixel. Matches cv2.resize bit for bit when shrinking
//      (tests/test_detector_gpu.py); OpenCV is third-party arithmetic (README.md:11).
__device__ __forceinline__ void lin_coef(int d, double scale, int ssize, int& s0, int& s1, int& a0, int& a1) {
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    int s = static_cast<int>(floorf(f));
    f = __fsub_rn(f, static_cast<float>(s));
    if (s < 0) {
        s = 0;
        f = 0.f;
    }
    if (s >= ssize - 1) {
        s = ssize - 1;
        f = 0.f;
    }
    s0 = s;
    s1 = min(s + 1, ssize - 1);
    a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}
__global__ void __launch_bounds__(256) det_letterbox_kernel(const uint8_t* __restrict__ frames, int frame_h, int frame_w, int stride,
                                                            int batch, int Hn, int Wn, int rw, int rh, int ox, int oy,
                                                            uint8_t* __restrict__ canvas) {
    griddep_launch_dependents();
    griddep_wait();
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(batch) * Hn * Wn) return;
    const int img = static_cast<int>(t / (Hn * Wn));
    const int rc = static_cast<int>(t - static_cast<long long>(img) * Hn * Wn);
    const int r = rc / Wn, c = rc % Wn;
    uint8_t* o = canvas + (static_cast<size_t>(img) * Hn * Wn + rc) * 3;
    const int dy = r - oy, dx = c - ox;
    if (dy < 0 || dy >= rh || dx < 0 || dx >= rw) {
        o[0] = o[1] = o[2] = 128;
        return;
    }
    int x0, x1, a0, a1, y0, y1, b0, b1;
    lin_coef(dx, static_cast<double>(frame_w) / rw, frame_w, x0, x1, a0, a1);
    lin_coef(dy, static_cast<double>(frame_h) / rh, frame_h, y0, y1, b0, b1);
    const uint8_t* base = frames + static_cast<size_t>(img) * frame_h * stride;
    const uint8_t* r0 = base + static_cast<size_t>(y0) * stride;
    const uint8_t* r1 = base + static_cast<size_t>(y1) * stride;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const int h0 = r0[x0 * 3 + ch] * a0 + r0[x1 * 3 + ch] * a1;
        const int h1 = r1[x0 * 3 + ch] * a0 + r1[x1 * 3 + ch] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[ch] = static_cast<uint8_t>(min(max(v, 0), 255));
    }
}

// ---- body.stage1.0: Conv3x3(3->8, s2, p1) + BN + ReLU on the letterboxed u8 canvas, fused with RetinaFace::preprocess's
//      convertTo(CV_32F) and mean subtraction (104, 117, 123) in B,G,R order (src/retinaface.cpp:128-135).
// The weights of the per-pixel kernels (this stem, the depthwise halves of dwpw_small_kernel) are KERNEL PARAMETERS: every lane of a
// warp multiplies by the same weight, so an FFMA takes it as a constant-bank operand (c[0x0][...]) and the shared-memory weight
// fetches (54 LDS.128 of 380 instructions per stem pixel) disappear. Same fp32 FMAs in the same order: results are bit-identical.
#ifndef FR_AB_CONSTW
#define FR_AB_CONSTW 1
#endif
struct StemW {
    float w[27][8];  // [tap * 3 + channel][output channel]
    float b[8];
};
struct DwW {
    float w[9][32];  // [tap][channel] (CIN <= 32)
    float b[32];
};
__global__ void __launch_bounds__(256) det_stem_kernel(const uint8_t* __restrict__ canvas, int stride_bytes, int batch, int Hn, int Wn,
                                                       const float* __restrict__ w /*[8][27]*/, const float* __restrict__ bias,
                                                       __half* __restrict__ out, const __grid_constant__ StemW cw) {
#if FR_AB_CONSTW
    const auto& ws = cw.w;
    const auto& sb = cw.b;
    griddep_launch_dependents();
    griddep_wait();
#else
    __shared__ float ws[27][8];
    __shared__ float sb[8];
    for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) ws[i % 27][i / 27] = w[i];
    if (threadIdx.x < 8) sb[threadIdx.x] = bias[threadIdx.x];
    griddep_launch_dependents();
    __syncthreads();
    griddep_wait();  // the weights staged above are static; the frames and the output map are not
#endif
    const Geo g{Hn / 2, Wn / 2};
    const int t = blockIdx.x * blockDim.x + threadIdx.x;  // pixels of a batch < 2^31 (checked by the host)
    if (t >= batch * g.H * g.W) return;
    const int img = t / (g.H * g.W);
    const int rc = t - img * (g.H * g.W);
    const int r = rc / g.W, c = rc % g.W;
    const float mean[3] = {104.f, 117.f, 123.f};
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = sb[n];
    const uint8_t* base = canvas + static_cast<size_t>(img) * Hn * stride_bytes;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int rr = 2 * r + ky - 1;
        if (rr < 0 || rr >= Hn) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int cc = 2 * c + kx - 1;
            if (cc < 0 || cc >= Wn) continue;
            const uint8_t* px = base + static_cast<size_t>(rr) * stride_bytes + cc * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float x = static_cast<float>(px[ch]) - mean[ch];
                const int k = (ky * 3 + kx) * 3 + ch;
#pragma unroll
                for (int n = 0; n < 8; ++n) acc[n] = fmaf(x, ws[k][n], acc[n]);
            }
        }
    }
    uint4 pk;
    __half2* hp = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) hp[j] = __floats2half2_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(img) * g.HpWp() + r * g.Wp() + c) * 8) = pk;
}

// same layer on the tensor RetinaFace::preprocess hands to the engine (f32 planar B,G,R, mean already subtracted): parity hook
__global__ void __launch_bounds__(256) det_stem_f32_kernel(const float* __restrict__ chw, int batch, int Hn, int Wn,
                                                           const float* __restrict__ w /*[8][27]*/, const float* __restrict__ bias,
                                                           __half* __restrict__ out) {
    __shared__ float ws[27][8];
    __shared__ float sb[8];
    for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) ws[i % 27][i / 27] = w[i];
    if (threadIdx.x < 8) sb[threadIdx.x] = bias[threadIdx.x];
    griddep_launch_dependents();
    __syncthreads();
    griddep_wait();  // the weights staged above are static; the frames and the output map are not
    const Geo g{Hn / 2, Wn / 2};
    const int t = blockIdx.x * blockDim.x + threadIdx.x;  // pixels of a batch < 2^31 (checked by the host)
    if (t >= batch * g.H * g.W) return;
    const int img = t / (g.H * g.W);
    const int rc = t - img * (g.H * g.W);
    const int r = rc / g.W, c = rc % g.W;
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = sb[n];
    const float* base = chw + static_cast<size_t>(img) * 3 * Hn * Wn;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int rr = 2 * r + ky - 1;
        if (rr < 0 || rr >= Hn) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int cc = 2 * c + kx - 1;
            if (cc < 0 || cc >= Wn) continue;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float x = base[(static_cast<size_t>(ch) * Hn + rr) * Wn + cc];
                const int k = (ky * 3 + kx) * 3 + ch;
#pragma unroll
                for (int n = 0; n < 8; ++n) acc[n] = fmaf(x, ws[k][n], acc[n]);
            }
        }
    }
    uint4 pk;
    __half2* hp = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) hp[j] = __floats2half2_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(img) * g.HpWp() + r * g.Wp() + c) * 8) = pk;
}

// ---- depthwise 3x3 (stride 1 or 2, pad 1) + BN + ReLU (first half of conv_dw, net.py:29-33) for the blocks with >= 64 channels
//      (their pointwise half is a tcgen05 GEMM). One work item = (output pixel, 8 channels); grid-stride, one full wave of CTAs, the
//      layer's weights ([9][C] f32, <= 9 KiB) and biases staged ONCE per CTA in shared memory. History (nine layers at 64 x 640^2):
//      weights fetched per item through L1 (two thirds of the load instructions) 382 us; the thread's 72 weights in registers
//      (118 registers, 16 warps per SM) 540 us - occupancy matters more than the fetches; this version keeps ~40 registers.
__global__ void __launch_bounds__(256) dw3x3_kernel(const __half* __restrict__ in, Geo gi, __half* __restrict__ out, Geo go, int stride,
                                                    int C, int batch, const float* __restrict__ w, const float* __restrict__ bias) {
    // weights as two planes of float4 per (tap, 8-channel chunk): channels 0-3 and 4-7 of the chunk, so that the lanes of a warp
    // (consecutive chunks) read consecutive 16-byte words - conflict-free LDS.128
    __shared__ float4 sw_lo[9 * 32], sw_hi[9 * 32];
    __shared__ __align__(16) float sbias[256];
    griddep_launch_dependents();
    for (int i = threadIdx.x; i < 9 * C; i += 256) {
        const int tap = i / C, cch = i - tap * C;
        float4* plane = (cch & 4) ? sw_hi : sw_lo;
        reinterpret_cast<float*>(&plane[tap * 32 + (cch >> 3)])[cch & 3] = __ldg(w + i);
    }
    for (int i = threadIdx.x; i < C; i += 256) sbias[i] = __ldg(bias + i);
    __syncthreads();
    griddep_wait();  // the weights are static; the input map belongs to the kernel before this one
    const int chunks = C / 8;                    // divides 256 (C = 64, 128, 256)
    const int ch = (threadIdx.x % chunks) * 8;   // this thread's channels: fixed, the grid stride is a multiple of `chunks`
    const int hw = go.H * go.W;
    const int pixels = batch * hw;               // < 2^31 (checked by the host)
    const int ppp = (gridDim.x * 256) / chunks;  // pixels the whole grid covers per pass
    for (int pix = (blockIdx.x * 256 + threadIdx.x) / chunks; pix < pixels; pix += ppp) {
        const int img = pix / hw;
        const int rc = pix - img * hw;
        const int r = rc / go.W, c = rc - r * go.W;
        float acc[8];
        {
            const float4 b0 = *reinterpret_cast<const float4*>(sbias + ch), b1 = *reinterpret_cast<const float4*>(sbias + ch + 4);
            acc[0] = b0.x, acc[1] = b0.y, acc[2] = b0.z, acc[3] = b0.w, acc[4] = b1.x, acc[5] = b1.y, acc[6] = b1.z, acc[7] = b1.w;
        }
        const __half* ibase = in + static_cast<size_t>(img) * gi.HpWp() * C + ch;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int rr = r * stride + ky - 1;
            if (rr < 0) continue;  // rr == gi.H is the zero pad row
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int cc = c * stride + kx - 1;
                if (cc < 0) continue;  // cc == gi.W is the zero pad column
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(ibase + (static_cast<size_t>(rr) * gi.Wp() + cc) * C));
                const __half2* h = reinterpret_cast<const __half2*>(&v);
                const float4 w0 = sw_lo[(ky * 3 + kx) * 32 + (ch >> 3)];
                const float4 w1 = sw_hi[(ky * 3 + kx) * 32 + (ch >> 3)];
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
        }
        uint4 pk;
        __half2* hp = reinterpret_cast<__half2*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) hp[j] = __floats2half2_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
        *reinterpret_cast<uint4*>(out + (static_cast<size_t>(img) * go.HpWp() + r * go.Wp() + c) * C + ch) = pk;
    }
}

// fp32 weights as mma.sync B fragments: {b0, b1} rounded to fp16 (x, y) and the fp16 of what the rounding lost (z, w)
__device__ __forceinline__ uint4 split_weight_frag(float w0, float w1, float w8, float w9) {
    const __half2 h0 = __floats2half2_rn(w0, w1), h1 = __floats2half2_rn(w8, w9);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(w0 - f0.x, w1 - f0.y), l1 = __floats2half2_rn(w8 - f1.x, w9 - f1.y);
    return make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1), *reinterpret_cast<const uint32_t*>(&l0),
                      *reinterpret_cast<const uint32_t*>(&l1));
}

// ---- FUSED conv_dw block for the early layers (net.py:29-38): depthwise 3x3 (stride S, pad 1) + BN + ReLU, then pointwise 1x1 +
//      BN + ReLU, CIN < 64. A warp owns 32 output pixels per pass:
//        1. depthwise: lane = pixel, the nine taps come straight from global memory through L1 (neighbouring pixels share them), fp32
//           accumulation in registers; the result (fp16, as the unfused pair of kernels stored it) goes to the warp's tile in shared
//           memory and never to global memory;
//        2. pointwise: the [32 pixels x CIN] tile times the [CIN x COUT] weights on the tensor cores with warp-level mma.sync
//           (m16n8k16 / m16n8k8 fragments via ldmatrix; the weight fragments are laid out once per CTA) instead of CIN * COUT FFMAs per
//           thread - the CUDA-core version was instruction-bound at 3x the time of its memory traffic;
//        3. bias + ReLU on the accumulator fragments, staged through shared memory and written back as whole rows (a lane-per-pixel
//           store would touch 32 different lines per instruction).
//      Grid-stride over groups of 32 pixels. dw_w: [9][CIN] f32, pw_w: [CIN][COUT] f32 (rounded to fp16 here, like every GEMM layer's).
template <int CIN, int COUT, int S>
__global__ void __launch_bounds__(256) dwpw_small_kernel(const __half* __restrict__ in, Geo gi, __half* __restrict__ out, Geo go, int batch,
                                                         const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                         const float* __restrict__ pw_w, const float* __restrict__ pw_b,
                                                         const __grid_constant__ DwW cw) {
    constexpr int KS = CIN >= 16 ? CIN / 16 : 1;  // k-steps (k16, or one k8 step for CIN = 8)
    constexpr int NT = COUT / 8;                  // 8-column output tiles
    constexpr int kInStride = CIN * 2 + 16;       // bytes per pixel row of the depthwise tile (padded: conflict-free ldmatrix)
    constexpr int kOutStride = COUT * 2 + 16;     // bytes per pixel row of the output tile
    constexpr int kTileBytes = 32 * (kInStride > kOutStride ? kInStride : kOutStride);
#if FR_AB_CONSTW
    const auto& sdw = cw.w;  // depthwise weights / bias as constant-bank operands (see StemW)
    const auto& sdb = cw.b;
#else
    __shared__ float sdw[9][CIN];
    __shared__ float sdb[CIN];
#endif
    __shared__ float spb[COUT];
    // weight fragments {b0, b1} of lane l for k-step ks, column tile nt, as fp16 hi + fp16 lo (w = hi + lo to ~22 bits: two MMAs per
    // fragment keep the fp32 weights' accuracy - with single fp16 weights the raw heads drifted past 1e-2 of the fp32 oracle)
    __shared__ uint4 sbf[KS][NT][32];
    __shared__ __align__(16) uint8_t tiles[8][kTileBytes];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t4 = lane & 3;
#if !FR_AB_CONSTW
    for (int i = threadIdx.x; i < 9 * CIN; i += blockDim.x) reinterpret_cast<float*>(&sdw[0][0])[i] = dw_w[i];
    for (int i = threadIdx.x; i < CIN; i += blockDim.x) sdb[i] = dw_b[i];
#endif
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) spb[i] = pw_b[i];
    for (int i = threadIdx.x; i < KS * NT * 32; i += blockDim.x) {
        const int l = i & 31, nt = (i >> 5) % NT, ks = i / (32 * NT);
        const int n = nt * 8 + (l >> 2), k0 = ks * 16 + (l & 3) * 2;
        auto w = [&](int k) { return k < CIN ? pw_w[k * COUT + n] : 0.f; };
        sbf[ks][nt][l] = split_weight_frag(w(k0), w(k0 + 1), w(k0 + 8), w(k0 + 9));
    }
    griddep_launch_dependents();
    __syncthreads();
    griddep_wait();  // weights are static; the input map belongs to the kernel before this one
    uint8_t* tile = tiles[warp];
    const uint32_t tile_addr = smem_u32(tile);
    const int hw = go.H * go.W;
    const int total = batch * hw;  // < 2^31 (checked by the host): 32-bit index arithmetic, no 64-bit divisions per pixel
    const int groups = (total + 31) / 32;
    for (int grp = blockIdx.x * 8 + warp; grp < groups; grp += gridDim.x * 8) {
        // ---------------- 1. depthwise, lane = pixel ----------------
        const int t = grp * 32 + lane;
        const bool live = t < total;
        int pos = 0;  // matrix row of this lane's output pixel
        {
            uint4 packed[CIN / 8];
            if (live) {
                const int img = t / hw;
                const int rc = t - img * hw;
                const int r = rc / go.W, c = rc - r * go.W;
                pos = img * go.HpWp() + r * go.Wp() + c;
                float x[CIN];
#pragma unroll
                for (int k = 0; k < CIN; ++k) x[k] = sdb[k];
                const __half* ibase = in + static_cast<size_t>(img) * gi.HpWp() * CIN;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int rr = r * S + ky - 1;
                    if (rr < 0) continue;  // rr == gi.H is the zero pad row
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int cc = c * S + kx - 1;
                        if (cc < 0) continue;  // cc == gi.W is the zero pad column
                        const uint4* src = reinterpret_cast<const uint4*>(ibase + (static_cast<size_t>(rr) * gi.Wp() + cc) * CIN);
                        uint4 v[CIN / 8];
                        if (CIN >= 16) {  // 32 bytes per load: half the L1 wavefronts of 16-byte loads at this stride
#pragma unroll
                            for (int ch = 0; ch < CIN / 8; ch += 2) ld_global_nc_256(src + ch, v[ch], v[ch + 1 < CIN / 8 ? ch + 1 : ch]);
                        } else {
                            v[0] = __ldg(src);
                        }
#pragma unroll
                        for (int ch = 0; ch < CIN / 8; ++ch) {
                            const __half2* h = reinterpret_cast<const __half2*>(&v[ch]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float2 f = __half22float2(h[q]);
                                x[ch * 8 + 2 * q] = fmaf(f.x, sdw[ky * 3 + kx][ch * 8 + 2 * q], x[ch * 8 + 2 * q]);
                                x[ch * 8 + 2 * q + 1] = fmaf(f.y, sdw[ky * 3 + kx][ch * 8 + 2 * q + 1], x[ch * 8 + 2 * q + 1]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int ch = 0; ch < CIN / 8; ++ch) {
                    __half2* hp = reinterpret_cast<__half2*>(&packed[ch]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) hp[q] = __floats2half2_rn(fmaxf(x[ch * 8 + 2 * q], 0.f), fmaxf(x[ch * 8 + 2 * q + 1], 0.f));
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < CIN / 8; ++ch) packed[ch] = make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int ch = 0; ch < CIN / 8; ++ch) *reinterpret_cast<uint4*>(tile + lane * kInStride + ch * 16) = packed[ch];
        }
        __syncwarp();
        // ---------------- 2. pointwise on the tensor cores: [32 x CIN] x [CIN x COUT] ----------------
        float acc[2][NT][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                if (CIN >= 16) {
                    // ldmatrix x4: lanes 0-7 / 8-15 / 16-23 / 24-31 address the rows of the (rows 0-7, k 0-7) / (rows 8-15, k 0-7) /
                    // (rows 0-7, k 8-15) / (rows 8-15, k 8-15) 8x8 blocks = fragments a0..a3
                    uint32_t a[4];
                    const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = ks * 16 + (lane >> 4) * 8;
                    ldmatrix_x4(a, tile_addr + row * kInStride + col * 2);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const uint4 b = sbf[ks][nt][lane];
                        mma_m16n8k16(acc[mt][nt], a, b.x, b.y);
                        mma_m16n8k16(acc[mt][nt], a, b.z, b.w);
                    }
                } else {
                    uint32_t a[2];
                    const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;  // lanes 16-31: addresses unused by x2 (kept valid)
                    ldmatrix_x2(a, tile_addr + row * kInStride);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const uint4 b = sbf[0][nt][lane];
                        mma_m16n8k8(acc[mt][nt], a, b.x);
                        mma_m16n8k8(acc[mt][nt], a, b.z);
                    }
                }
            }
        }
        __syncwarp();  // every lane has read its A fragments: the tile may be overwritten with the outputs
        // ---------------- 3. bias + ReLU, staged rows, whole-row stores ----------------
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = nt * 8 + t4 * 2;
                const float b0 = spb[col], b1 = spb[col + 1];
                const __half2 lo = __floats2half2_rn(fmaxf(acc[mt][nt][0] + b0, 0.f), fmaxf(acc[mt][nt][1] + b1, 0.f));
                const __half2 hi = __floats2half2_rn(fmaxf(acc[mt][nt][2] + b0, 0.f), fmaxf(acc[mt][nt][3] + b1, 0.f));
                *reinterpret_cast<__half2*>(tile + (mt * 16 + g) * kOutStride + col * 2) = lo;
                *reinterpret_cast<__half2*>(tile + (mt * 16 + g + 8) * kOutStride + col * 2) = hi;
            }
        __syncwarp();
        constexpr int CPR = COUT * 2 / 16;  // 16-byte chunks per output row
        constexpr int RPI = 32 / CPR;       // rows per store instruction
        const unsigned live_mask = __ballot_sync(0xffffffffu, live);
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int q = i * RPI + lane / CPR, chunk = lane % CPR;
            const int pos_q = __shfl_sync(0xffffffffu, pos, q);
            const uint4 v = *reinterpret_cast<const uint4*>(tile + q * kOutStride + chunk * 16);
            if ((live_mask >> q) & 1u) *reinterpret_cast<uint4*>(out + static_cast<size_t>(pos_q) * COUT + chunk * 8) = v;
        }
        __syncwarp();  // the tile is rewritten by the next pass
    }
}

// ---- SSH 16 -> 16 3x3 conv + BN + ReLU (conv5X5_2, conv7X7_2, conv7x7_3, net.py:49-53; the ReLU is either the layer's own or
//      the one applied to the concatenation, net.py:64-65). Output may be a channel slice of a wider map (ld_out, pre-offset pointer).
//      NCONV = 2: conv5X5_2 and conv7X7_2 read the same map (net.py:58-61), so one pass over it computes both (second weight set wB,
//      second destination outB). w: [9][16][16] f32 (tap, cin, cout), rounded to fp16 here.
//      A warp owns 32 pixels per pass; per kernel row ky every lane stages its pixel's three taps (3 x 32 bytes, zero outside the map)
//      in the warp's shared-memory tile, and the [32 pixels x 16 channels] x [16 x 16*NCONV] product of each tap runs on the tensor
//      cores (mma.sync m16n8k16, fragments via ldmatrix): 36 / 72 MMAs per pass instead of 2304 / 4608 FFMAs per thread.
template <int NCONV>
__global__ void __launch_bounds__(NCONV == 2 ? 128 : 256) conv3x3_c16_kernel(const __half* __restrict__ in, Geo g, int batch, const float* __restrict__ wA,
                                                          const float* __restrict__ bA, __half* __restrict__ outA, int ldA,
                                                          const float* __restrict__ wB, const float* __restrict__ bB,
                                                          __half* __restrict__ outB, int ldB) {
    constexpr int NT = 2 * NCONV;   // 8-column output tiles
    constexpr int kRow = 32 + 16;   // bytes per staged pixel (16 channels + pad: conflict-free ldmatrix)
    constexpr int kWarps = NCONV == 2 ? 4 : 8;  // 48 KiB of static shared memory: fragments (hi + lo) + one tile per warp
    __shared__ uint4 sbf[9][NT][32];  // weight fragments {b0, b1} as fp16 hi + lo (split_weight_frag) per tap / column tile / lane
    __shared__ float sb[16 * NCONV];
    __shared__ __align__(16) uint8_t tiles[kWarps][3 * 32 * kRow];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, t4 = lane & 3;
    for (int i = threadIdx.x; i < 9 * NT * 32; i += blockDim.x) {
        const int l = i & 31, nt = (i >> 5) % NT, tap = i / (32 * NT);
        const float* w = nt < 2 ? wA : wB;
        const int n = (nt & 1) * 8 + (l >> 2), k0 = (l & 3) * 2;
        auto wv = [&](int k) { return w[(tap * 16 + k) * 16 + n]; };
        sbf[tap][nt][l] = split_weight_frag(wv(k0), wv(k0 + 1), wv(k0 + 8), wv(k0 + 9));
    }
    if (threadIdx.x < 16 * NCONV) sb[threadIdx.x] = threadIdx.x < 16 ? bA[threadIdx.x] : bB[threadIdx.x - 16];
    griddep_launch_dependents();
    __syncthreads();
    griddep_wait();
    uint8_t* tile = tiles[warp];
    const uint32_t tile_addr = smem_u32(tile);
    const int hw = g.H * g.W;
    const int total = batch * hw;  // < 2^31 (checked by the host)
    const int groups = (total + 31) / 32;
    for (int grp = blockIdx.x * kWarps + warp; grp < groups; grp += gridDim.x * kWarps) {
        const int t = grp * 32 + lane;
        const bool live = t < total;
        int img = 0, r = 0, c = 0, pos = 0;
        if (live) {
            img = t / hw;
            const int rc = t - img * hw;
            r = rc / g.W;
            c = rc - r * g.W;
            pos = img * g.HpWp() + r * g.Wp() + c;
        }
        float acc[2][NT][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;
        const __half* ibase = in + static_cast<size_t>(img) * g.HpWp() * 16;
#pragma unroll 1
        for (int ky = 0; ky < 3; ++ky) {
            // stage this lane's three taps of kernel row ky (row g.H / column g.W are the layout's zero pads, read as stored)
            const int rr = r + ky - 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int cc = c + kx - 1;
                uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
                if (live && rr >= 0 && cc >= 0) ld_global_nc_256(ibase + (static_cast<size_t>(rr) * g.Wp() + cc) * 16, v0, v1);
                uint4* dst = reinterpret_cast<uint4*>(tile + (kx * 32 + lane) * kRow);
                dst[0] = v0;
                dst[1] = v1;
            }
            __syncwarp();
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    uint32_t a[4];
                    const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = (lane >> 4) * 8;
                    ldmatrix_x4(a, tile_addr + (kx * 32 + row) * kRow + col * 2);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const uint4 b = sbf[ky * 3 + kx][nt][lane];
                        mma_m16n8k16(acc[mt][nt], a, b.x, b.y);
                        mma_m16n8k16(acc[mt][nt], a, b.z, b.w);
                    }
                }
            __syncwarp();  // all fragments are in registers: the tile may be restaged
        }
        // bias + ReLU -> staged rows [32 pixels][16 * NCONV channels] -> row stores (conv A: channels 0-15, conv B: 16-31)
        constexpr int kOut = 32 * NCONV + 16;  // bytes per staged output row
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = nt * 8 + t4 * 2;
                const float b0 = sb[col], b1 = sb[col + 1];
                const __half2 lo = __floats2half2_rn(fmaxf(acc[mt][nt][0] + b0, 0.f), fmaxf(acc[mt][nt][1] + b1, 0.f));
                const __half2 hi = __floats2half2_rn(fmaxf(acc[mt][nt][2] + b0, 0.f), fmaxf(acc[mt][nt][3] + b1, 0.f));
                *reinterpret_cast<__half2*>(tile + (mt * 16 + gq) * kOut + col * 2) = lo;
                *reinterpret_cast<__half2*>(tile + (mt * 16 + gq + 8) * kOut + col * 2) = hi;
            }
        __syncwarp();
        const unsigned live_mask = __ballot_sync(0xffffffffu, live);
#pragma unroll
        for (int cv = 0; cv < NCONV; ++cv) {
            __half* o = cv == 0 ? outA : outB;
            const int ld = cv == 0 ? ldA : ldB;
#pragma unroll
            for (int i = 0; i < 2; ++i) {  // 16 pixels x 2 chunks per instruction
                const int q = i * 16 + (lane >> 1), chunk = lane & 1;
                const int pos_q = __shfl_sync(0xffffffffu, pos, q);
                const uint4 v = *reinterpret_cast<const uint4*>(tile + q * kOut + cv * 32 + chunk * 16);
                if ((live_mask >> q) & 1u) *reinterpret_cast<uint4*>(o + static_cast<size_t>(pos_q) * ld + chunk * 8) = v;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// RetinaFace::postprocessing + create_anchor_retinaface + nms (src/retinaface.cpp:154-271), one block per image.
// The arithmetic types follow the C++ statement by statement (double intermediates, float stores, two truncations to int);
// explicit _rn intrinsics keep the compiler from contracting multiply-adds the host code does not contract.
// Greedy NMS over the score-sorted list is evaluated as: repeat { take the best live candidate (score desc, anchor asc);
// kill every live candidate whose IoU with it is >= thr } — the same survivors, in the same order, as the reference's
// erase loop; it stops after max_faces survivors because later candidates cannot change earlier ones (:206-207).
// ---------------------------------------------------------------------------------------------------------------
struct DetPostParams {
    int net_w, net_h, frame_w, frame_h;
    float nms_thr, bbox_thr;
    int max_faces;
    int anchors;
};

struct DetCand {
    int x1, y1, x2, y2;
    float score;
    int id;
};

__device__ __forceinline__ void anchor_of(int id, int net_w, int net_h, float& cx, float& cy, float& sx, float& sy) {
    const float steps[3] = {8.f, 16.f, 32.f};
    const int min_sizes[3][2] = {{10, 20}, {32, 64}, {128, 256}};
    int k = 0, base = 0;
    int fw = 0;
    for (; k < 3; ++k) {
        const int fh = static_cast<int>(ceilf(net_h / steps[k]));
        fw = static_cast<int>(ceilf(net_w / steps[k]));
        const int cnt = fh * fw * 2;
        if (id < base + cnt || k == 2) break;
        base += cnt;
    }
    const int rel = id - base;
    const int l = rel & 1, cell = rel >> 1;
    const int i = cell / fw, j = cell - i * fw;
    sx = static_cast<float>(__ddiv_rn(static_cast<double>(min_sizes[k][l]) * 1.0, static_cast<double>(net_w)));
    sy = static_cast<float>(__ddiv_rn(static_cast<double>(min_sizes[k][l]) * 1.0, static_cast<double>(net_h)));
    cx = static_cast<float>(__ddiv_rn(__dmul_rn(static_cast<double>(j) + 0.5, static_cast<double>(steps[k])), static_cast<double>(net_w)));
    cy = static_cast<float>(__ddiv_rn(__dmul_rn(static_cast<double>(i) + 0.5, static_cast<double>(steps[k])), static_cast<double>(net_h)));
}

__device__ __forceinline__ int clipi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }

// decode of one anchor that passed the threshold (src/retinaface.cpp:163-200)
__device__ __forceinline__ DetCand decode_candidate(const float4 b, float score, int a, const DetPostParams& prm, float scale_h, float scale_w) {
    float cx, cy, sx, sy;
    anchor_of(a, prm.net_w, prm.net_h, cx, cy, sx, sy);
    const float tcx = static_cast<float>(__dadd_rn(static_cast<double>(cx), __dmul_rn(__dmul_rn(static_cast<double>(b.x), 0.1), static_cast<double>(sx))));
    const float tcy = static_cast<float>(__dadd_rn(static_cast<double>(cy), __dmul_rn(__dmul_rn(static_cast<double>(b.y), 0.1), static_cast<double>(sy))));
    const float tsx = static_cast<float>(__dmul_rn(static_cast<double>(sx), exp(__dmul_rn(static_cast<double>(b.z), 0.2))));
    const float tsy = static_cast<float>(__dmul_rn(static_cast<double>(sy), exp(__dmul_rn(static_cast<double>(b.w), 0.2))));
    const float hx = __fdiv_rn(tsx, 2.f), hy = __fdiv_rn(tsy, 2.f);
    int y1 = static_cast<int>(__fmul_rn(__fsub_rn(tcx, hx), static_cast<float>(prm.net_w)));  // :171-174
    int x1 = static_cast<int>(__fmul_rn(__fsub_rn(tcy, hy), static_cast<float>(prm.net_h)));
    int y2 = static_cast<int>(__fmul_rn(__fadd_rn(tcx, hx), static_cast<float>(prm.net_w)));
    int x2 = static_cast<int>(__fmul_rn(__fadd_rn(tcy, hy), static_cast<float>(prm.net_h)));
    if (scale_h > scale_w) {  // :177-187
        const float pad = __fdiv_rn(__fsub_rn(static_cast<float>(prm.net_h), __fmul_rn(scale_w, static_cast<float>(prm.frame_h))), 2.f);
        y1 = static_cast<int>(__fdiv_rn(static_cast<float>(y1), scale_w));
        y2 = static_cast<int>(__fdiv_rn(static_cast<float>(y2), scale_w));
        x1 = static_cast<int>(__fdiv_rn(__fsub_rn(static_cast<float>(x1), pad), scale_w));
        x2 = static_cast<int>(__fdiv_rn(__fsub_rn(static_cast<float>(x2), pad), scale_w));
    } else {
        const float pad = __fdiv_rn(__fsub_rn(static_cast<float>(prm.net_w), __fmul_rn(scale_h, static_cast<float>(prm.frame_w))), 2.f);
        y1 = static_cast<int>(__fdiv_rn(__fsub_rn(static_cast<float>(y1), pad), scale_h));
        y2 = static_cast<int>(__fdiv_rn(__fsub_rn(static_cast<float>(y2), pad), scale_h));
        x1 = static_cast<int>(__fdiv_rn(static_cast<float>(x1), scale_h));
        x2 = static_cast<int>(__fdiv_rn(static_cast<float>(x2), scale_h));
    }
    DetCand d;
    d.y1 = clipi(y1, 0, prm.frame_w - 1);  // :190-193
    d.x1 = clipi(x1, 0, prm.frame_h - 1);
    d.y2 = clipi(y2, 0, prm.frame_w - 1);
    d.x2 = clipi(x2, 0, prm.frame_h - 1);
    d.score = score;
    d.id = a;
    return d;
}

// stage 1, grid (ceil(anchors / 256), batch): threshold + decode, one thread per anchor; survivors are appended to the image's
// candidate list (order irrelevant: the NMS below picks by (score, anchor id)). n_cand[img] must be zero on entry; det_nms_kernel
// leaves it zero again.
__global__ void __launch_bounds__(256) det_decode_kernel(const float* __restrict__ loc, const float* __restrict__ conf, DetPostParams prm,
                                                         DetCand* __restrict__ ws, int* __restrict__ n_cand) {
    griddep_launch_dependents();
    griddep_wait();
    const int img = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= prm.anchors) return;
    const float score = conf[(static_cast<size_t>(img) * prm.anchors + a) * 2 + 1];
    if (!(score > prm.bbox_thr)) return;  // strict >, :160
    const float scale_h = __fdiv_rn(static_cast<float>(prm.net_h), static_cast<float>(prm.frame_h));  // :21
    const float scale_w = __fdiv_rn(static_cast<float>(prm.net_w), static_cast<float>(prm.frame_w));  // :22
    const float4 b = *reinterpret_cast<const float4*>(loc + (static_cast<size_t>(img) * prm.anchors + a) * 4);
    ws[static_cast<size_t>(img) * prm.anchors + atomicAdd(&n_cand[img], 1)] = decode_candidate(b, score, a, prm, scale_h, scale_w);
}

// stage 2, one block per image: greedy NMS over the image's candidates
__global__ void __launch_bounds__(256) det_nms_kernel(const float* __restrict__ landm, DetPostParams prm, DetCand* __restrict__ ws,
                                                      int* __restrict__ n_cand_g, FrBbox* __restrict__ boxes, int* __restrict__ counts,
                                                      float* __restrict__ out_landm, int* __restrict__ out_ids) {
    __shared__ float red_s[8];
    __shared__ int red_i[8], red_p[8];
    __shared__ DetCand kept;
    __shared__ int kept_pos;
    __shared__ int n_cand;
    griddep_launch_dependents();
    griddep_wait();
    const int img = blockIdx.x;
    DetCand* cand = ws + static_cast<size_t>(img) * prm.anchors;
    const float scale_h = __fdiv_rn(static_cast<float>(prm.net_h), static_cast<float>(prm.frame_h));  // :21
    const float scale_w = __fdiv_rn(static_cast<float>(prm.net_w), static_cast<float>(prm.frame_w));  // :22
    if (threadIdx.x == 0) {
        n_cand = n_cand_g[img];
        n_cand_g[img] = 0;  // ready for the next batch
    }
    __syncthreads();
    const int n = n_cand;
    int n_kept = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (n_kept < prm.max_faces) {
        // best live candidate: highest score, lowest anchor id among equals (a dead candidate has id < 0)
        float bs = -1.f;
        int bi = 0x7fffffff, bp = -1;
        for (int p = threadIdx.x; p < n; p += blockDim.x) {
            const int id = cand[p].id;
            if (id < 0) continue;
            const float s = cand[p].score;
            if (bp < 0 || s > bs || (s == bs && id < bi)) {
                bs = s;
                bi = id;
                bp = p;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (op >= 0 && (bp < 0 || os > bs || (os == bs && oi < bi))) {
                bs = os;
                bi = oi;
                bp = op;
            }
        }
        if (lane == 0) {
            red_s[warp] = bs;
            red_i[warp] = bi;
            red_p[warp] = bp;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w2 = 1; w2 < 8; ++w2)
                if (red_p[w2] >= 0 && (bp < 0 || red_s[w2] > bs || (red_s[w2] == bs && red_i[w2] < bi))) {
                    bs = red_s[w2];
                    bi = red_i[w2];
                    bp = red_p[w2];
                }
            kept_pos = bp;
            if (bp >= 0) kept = cand[bp];
        }
        __syncthreads();
        if (kept_pos < 0) break;
        const DetCand k = kept;
        if (threadIdx.x == 0) {
            FrBbox b;
            b.x1 = k.x1;
            b.y1 = k.y1;
            b.x2 = k.x2;
            b.y2 = k.y2;
            b.score = k.score;
            boxes[static_cast<size_t>(img) * prm.max_faces + n_kept] = b;
            if (out_ids) out_ids[static_cast<size_t>(img) * prm.max_faces + n_kept] = k.id;
            if (out_landm) {
                float* lm = out_landm + (static_cast<size_t>(img) * prm.max_faces + n_kept) * 10;
                if (landm) {
                    float cx, cy, sx, sy;
                    anchor_of(k.id, prm.net_w, prm.net_h, cx, cy, sx, sy);
                    const float* l = landm + (static_cast<size_t>(img) * prm.anchors + k.id) * 10;
                    for (int pt = 0; pt < 5; ++pt) {
                        const float lx = static_cast<float>(__dadd_rn(static_cast<double>(cx), __dmul_rn(__dmul_rn(static_cast<double>(l[2 * pt]), 0.1), static_cast<double>(sx))));
                        const float ly = static_cast<float>(__dadd_rn(static_cast<double>(cy), __dmul_rn(__dmul_rn(static_cast<double>(l[2 * pt + 1]), 0.1), static_cast<double>(sy))));
                        float px = __fmul_rn(lx, static_cast<float>(prm.net_w)), py = __fmul_rn(ly, static_cast<float>(prm.net_h));
                        if (scale_h > scale_w) {
                            const float pad = __fdiv_rn(__fsub_rn(static_cast<float>(prm.net_h), __fmul_rn(scale_w, static_cast<float>(prm.frame_h))), 2.f);
                            px = __fdiv_rn(px, scale_w);
                            py = __fdiv_rn(__fsub_rn(py, pad), scale_w);
                        } else {
                            const float pad = __fdiv_rn(__fsub_rn(static_cast<float>(prm.net_w), __fmul_rn(scale_h, static_cast<float>(prm.frame_w))), 2.f);
                            px = __fdiv_rn(__fsub_rn(px, pad), scale_h);
                            py = __fdiv_rn(py, scale_h);
                        }
                        lm[2 * pt] = px;
                        lm[2 * pt + 1] = py;
                    }
                } else {
                    for (int q = 0; q < 10; ++q) lm[q] = 0.f;
                }
            }
        }
        // suppression pass: IoU with the +1 area convention and '>=' (:251,259-263)
        const float area_k = static_cast<float>((k.x2 - k.x1 + 1) * (k.y2 - k.y1 + 1));
        for (int p = threadIdx.x; p < n; p += blockDim.x) {
            DetCand d = cand[p];
            if (d.id < 0) continue;
            bool kill = (p == kept_pos);
            if (!kill) {
                const float xx1 = static_cast<float>(max(k.x1, d.x1)), yy1 = static_cast<float>(max(k.y1, d.y1));
                const float xx2 = static_cast<float>(min(k.x2, d.x2)), yy2 = static_cast<float>(min(k.y2, d.y2));
                const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
                const float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
                const float inter = __fmul_rn(w, h);
                const float area_d = static_cast<float>((d.x2 - d.x1 + 1) * (d.y2 - d.y1 + 1));
                const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_k, area_d), inter));
                kill = ovr >= prm.nms_thr;
            }
            if (kill) cand[p].id = -1;
        }
        ++n_kept;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[img] = n_kept;
}

}  // namespace frb
