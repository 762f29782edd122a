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
__global__ void __launch_bounds__(256) det_stem_kernel(const uint8_t* __restrict__ canvas, int stride_bytes, int batch, int Hn, int Wn,
                                                       const float* __restrict__ w /*[8][27]*/, const float* __restrict__ bias,
                                                       __half* __restrict__ out) {
    __shared__ float ws[27][8];
    __shared__ float sb[8];
    for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) ws[i % 27][i / 27] = w[i];
    if (threadIdx.x < 8) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const Geo g{Hn / 2, Wn / 2};
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(batch) * g.H * g.W) return;
    const int img = static_cast<int>(t / (g.H * g.W));
    const int rc = static_cast<int>(t - static_cast<long long>(img) * g.H * g.W);
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
    __syncthreads();
    const Geo g{Hn / 2, Wn / 2};
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(batch) * g.H * g.W) return;
    const int img = static_cast<int>(t / (g.H * g.W));
    const int rc = static_cast<int>(t - static_cast<long long>(img) * g.H * g.W);
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

// ---- depthwise 3x3 (stride 1 or 2, pad 1) + BN + ReLU (first half of conv_dw, net.py:29-33). One thread per
//      (output pixel, 8 channels). w: [9][C] f32.
__global__ void __launch_bounds__(256) dw3x3_kernel(const __half* __restrict__ in, Geo gi, __half* __restrict__ out, Geo go, int stride,
                                                    int C, int batch, const float* __restrict__ w, const float* __restrict__ bias) {
    const int chunks = C / 8;
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(batch) * go.H * go.W * chunks) return;
    const int ch = static_cast<int>(t % chunks) * 8;
    const long long pix = t / chunks;
    const int img = static_cast<int>(pix / (go.H * go.W));
    const int rc = static_cast<int>(pix - static_cast<long long>(img) * go.H * go.W);
    const int r = rc / go.W, c = rc % go.W;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + ch + j);
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
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + ch));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + ch) + 1);
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

// ---- FUSED conv_dw block for the early layers (net.py:29-38): depthwise 3x3 (stride S, pad 1) + BN + ReLU, then pointwise 1x1 +
//      BN + ReLU, CIN < 64. One thread per output pixel: the nine taps come straight from global memory through L1 (neighbouring
//      pixels share them), the depthwise result stays in registers (fp32, never written to global memory) and feeds the CIN -> COUT
//      pointwise product, whose weights are broadcast from shared memory. Grid-stride: the weights are staged once per CTA.
//      Compared with dw3x3_kernel + pw_small_kernel this removes the write + re-read of the depthwise map and one launch per block.
//      dw_w: [9][CIN] f32, pw_w: [CIN][COUT] f32.
template <int CIN, int COUT, int S>
__global__ void __launch_bounds__(256) dwpw_small_kernel(const __half* __restrict__ in, Geo gi, __half* __restrict__ out, Geo go, int batch,
                                                         const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                         const float* __restrict__ pw_w, const float* __restrict__ pw_b) {
    __shared__ float sdw[9][CIN];
    __shared__ float sdb[CIN];
    __shared__ float4 spw[CIN][COUT / 4];
    __shared__ float spb[COUT];
    for (int i = threadIdx.x; i < 9 * CIN; i += blockDim.x) reinterpret_cast<float*>(&sdw[0][0])[i] = dw_w[i];
    for (int i = threadIdx.x; i < CIN * COUT; i += blockDim.x) reinterpret_cast<float*>(&spw[0][0])[i] = pw_w[i];
    for (int i = threadIdx.x; i < CIN; i += blockDim.x) sdb[i] = dw_b[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) spb[i] = pw_b[i];
    __syncthreads();
    const long long total = static_cast<long long>(batch) * go.H * go.W;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int img = static_cast<int>(t / (go.H * go.W));
        const int rc = static_cast<int>(t - static_cast<long long>(img) * go.H * go.W);
        const int r = rc / go.W, c = rc - r * go.W;
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
#pragma unroll
                for (int ch = 0; ch < CIN / 8; ++ch) {
                    const uint4 v = __ldg(src + ch);
                    const __half2* h = reinterpret_cast<const __half2*>(&v);
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
        for (int k = 0; k < CIN; ++k) x[k] = fmaxf(x[k], 0.f);
        __half* dst_px = out + (static_cast<size_t>(img) * go.HpWp() + static_cast<size_t>(r) * go.Wp() + c) * COUT;
#pragma unroll 1
        for (int n0 = 0; n0 < COUT; n0 += 16) {
            float acc[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = spb[n0 + q];
#pragma unroll
            for (int k = 0; k < CIN; ++k) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 wv = spw[k][n0 / 4 + j4];
                    acc[4 * j4 + 0] = fmaf(x[k], wv.x, acc[4 * j4 + 0]);
                    acc[4 * j4 + 1] = fmaf(x[k], wv.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(x[k], wv.z, acc[4 * j4 + 2]);
                    acc[4 * j4 + 3] = fmaf(x[k], wv.w, acc[4 * j4 + 3]);
                }
            }
            uint4 pk[2];
            __half2* hp = reinterpret_cast<__half2*>(pk);
#pragma unroll
            for (int q = 0; q < 8; ++q) hp[q] = __floats2half2_rn(fmaxf(acc[2 * q], 0.f), fmaxf(acc[2 * q + 1], 0.f));
            st_global_256(dst_px + n0, pk[0], pk[1]);
        }
    }
}

// ---- SSH 16 -> 16 3x3 conv + BN + ReLU (conv5X5_2, conv7X7_2, conv7x7_3, net.py:49-53; the ReLU is either the layer's own or
//      the one applied to the concatenation, net.py:64-65). w: [9][16][16] f32 (tap, cin, cout). Output may be a channel slice
//      of a wider map (ld_out, pre-offset pointer). NCONV = 2: conv5X5_2 and conv7X7_2 read the same map (net.py:58-61), so one
//      pass over it computes both (second weight set wB, second destination outB).
template <int NCONV>
__global__ void __launch_bounds__(128) conv3x3_c16_kernel(const __half* __restrict__ in, Geo g, int batch, const float* __restrict__ wA,
                                                          const float* __restrict__ bA, __half* __restrict__ outA, int ldA,
                                                          const float* __restrict__ wB, const float* __restrict__ bB,
                                                          __half* __restrict__ outB, int ldB) {
    constexpr int NO = 16 * NCONV;
    __shared__ float4 ws[9 * 16][NO / 4];
    __shared__ float sb[NO];
    for (int i = threadIdx.x; i < 9 * 16 * 16; i += blockDim.x) {
        reinterpret_cast<float*>(&ws[i / 16][0])[i % 16] = wA[i];
        if (NCONV == 2) reinterpret_cast<float*>(&ws[i / 16][0])[16 + i % 16] = wB[i];
    }
    if (threadIdx.x < 16) {
        sb[threadIdx.x] = bA[threadIdx.x];
        if (NCONV == 2) sb[16 + threadIdx.x] = bB[threadIdx.x];
    }
    __syncthreads();
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(batch) * g.H * g.W) return;
    const int img = static_cast<int>(t / (g.H * g.W));
    const int rc = static_cast<int>(t - static_cast<long long>(img) * g.H * g.W);
    const int r = rc / g.W, c = rc % g.W;
    float acc[NO];
#pragma unroll
    for (int j = 0; j < NO; ++j) acc[j] = sb[j];
    const __half* ibase = in + static_cast<size_t>(img) * g.HpWp() * 16;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int rr = r + ky - 1;
        if (rr < 0) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int cc = c + kx - 1;
            if (cc < 0) continue;
            const uint4* src = reinterpret_cast<const uint4*>(ibase + (static_cast<size_t>(rr) * g.Wp() + cc) * 16);
            const uint4 v0 = __ldg(src), v1 = __ldg(src + 1);
            float x[16];
            const __half2* h0 = reinterpret_cast<const __half2*>(&v0);
            const __half2* h1 = reinterpret_cast<const __half2*>(&v1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 a = __half22float2(h0[j]), b = __half22float2(h1[j]);
                x[2 * j] = a.x;
                x[2 * j + 1] = a.y;
                x[8 + 2 * j] = b.x;
                x[8 + 2 * j + 1] = b.y;
            }
            const int tap = ky * 3 + kx;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
#pragma unroll
                for (int j4 = 0; j4 < NO / 4; ++j4) {
                    const float4 wv = ws[tap * 16 + k][j4];
                    acc[4 * j4 + 0] = fmaf(x[k], wv.x, acc[4 * j4 + 0]);
                    acc[4 * j4 + 1] = fmaf(x[k], wv.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(x[k], wv.z, acc[4 * j4 + 2]);
                    acc[4 * j4 + 3] = fmaf(x[k], wv.w, acc[4 * j4 + 3]);
                }
            }
        }
    }
    const size_t pos = static_cast<size_t>(img) * g.HpWp() + r * g.Wp() + c;
#pragma unroll
    for (int q = 0; q < NCONV; ++q) {
        uint4 pk[2];
        __half2* hp = reinterpret_cast<__half2*>(pk);
#pragma unroll
        for (int j = 0; j < 8; ++j) hp[j] = __floats2half2_rn(fmaxf(acc[q * 16 + 2 * j], 0.f), fmaxf(acc[q * 16 + 2 * j + 1], 0.f));
        st_global_256(q == 0 ? outA + pos * ldA : outB + pos * ldB, pk[0], pk[1]);
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
