// CUDA-core kernels around the tensor-core conv GEMM of the ArcFace IR-(SE)50 embedder
// (network spec: /root/reference conversion/arcface/model_irse.py; preprocessing: src/arcface.cpp:105-129).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace frb {

// ---------------------------------------------------------------------------------------------------------------
// input_layer: Conv3x3(3->64, s1, p1) + BN (folded) + PReLU  (model_irse.py:139-141). Cin = 3 is not GEMM-shaped: direct conv.
// Input: either the tensor ArcFaceIR50::preprocessFaces produces (f32 planar R,G,B, (x-127.5)*0.0078125, src/arcface.cpp:118-125)
// or the u8 BGR HWC crop itself (the same arithmetic is applied on the fly).
// Output (shared-halo flat NHWC, H = W = 112): y and y_bn = y * bn_s + bn_b (unit 0's pre-activation BN).
// One thread per PAIR of horizontally adjacent pixels, 64 output channels in four groups of 16: every weight vector fetched from
// shared memory (LDS.128, broadcast) feeds eight FMAs instead of four, and the pair shares 6 of its 9 input columns. Each output is
// the same bias + 27-term FMA chain in the same order as a one-pixel-per-thread kernel (bit-identical results).
// ---------------------------------------------------------------------------------------------------------------
template <bool kU8>
__global__ void __launch_bounds__(128) arcface_stem_pair_kernel(const void* __restrict__ in, int batch, const float* __restrict__ w /*[64][27]*/,
                                                           const float* __restrict__ bias, const float* __restrict__ prelu,
                                                           const float* __restrict__ bn_s, const float* __restrict__ bn_b,
                                                           __half* __restrict__ y, __half* __restrict__ y_bn) {
    constexpr int S = 112, Wp = S + 1, HpWp = Wp * Wp, kPairsPerRow = S / 2;
    __shared__ float4 ws[27][16];  // ws[k][n/4] = w[n..n+3][k]
    __shared__ float sb[64], sp[64], ss[64], sbb[64];
    for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
        const int k = i / 64, n = i % 64;
        reinterpret_cast<float*>(&ws[k][0])[n] = w[n * 27 + k];
    }
    if (threadIdx.x < 64) {
        sb[threadIdx.x] = bias[threadIdx.x];
        sp[threadIdx.x] = prelu[threadIdx.x];
        ss[threadIdx.x] = bn_s ? bn_s[threadIdx.x] : 1.f;
        sbb[threadIdx.x] = bn_b ? bn_b[threadIdx.x] : 0.f;
    }
    __syncthreads();
    // grid-stride over the pixel pairs: the 6.9 KiB of weights staged above are amortised over several pairs per thread
    const long long total = static_cast<long long>(batch) * S * kPairsPerRow;
    for (long long pr = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pr < total;
         pr += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int img = static_cast<int>(pr / (S * kPairsPerRow));
        const int rc = static_cast<int>(pr - static_cast<long long>(img) * S * kPairsPerRow);
        const int r = rc / kPairsPerRow, c = (rc % kPairsPerRow) * 2;
        float xin[3][4][3];  // [ky][input column c - 1 + j][ch], ch in R,G,B order
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int rr = r + ky - 1, cc = c + j - 1;
                const bool ok = rr >= 0 && rr < S && cc >= 0 && cc < S;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float v = 0.f;
                    if (ok) {
                        if (kU8) {
                            const uint8_t u = static_cast<const uint8_t*>(in)[(static_cast<size_t>(img) * S * S + rr * S + cc) * 3 + (2 - ch)];
                            v = (static_cast<float>(u) - 127.5f) * 0.0078125f;
                        } else {
                            v = static_cast<const float*>(in)[(static_cast<size_t>(img) * 3 + ch) * S * S + rr * S + cc];
                        }
                    }
                    xin[ky][j][ch] = v;
                }
            }
        const size_t o = (static_cast<size_t>(img) * HpWp + r * Wp + c) * 64;  // pixel c; pixel c + 1 follows 64 channels later
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
            float acc[2][16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[0][j] = acc[1][j] = sb[g * 16 + j];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const int k = (ky * 3 + kx) * 3 + ch;
                        const float a0 = xin[ky][kx][ch], a1 = xin[ky][kx + 1][ch];
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 wv = ws[k][g * 4 + j4];
                            acc[0][4 * j4 + 0] = fmaf(a0, wv.x, acc[0][4 * j4 + 0]);
                            acc[0][4 * j4 + 1] = fmaf(a0, wv.y, acc[0][4 * j4 + 1]);
                            acc[0][4 * j4 + 2] = fmaf(a0, wv.z, acc[0][4 * j4 + 2]);
                            acc[0][4 * j4 + 3] = fmaf(a0, wv.w, acc[0][4 * j4 + 3]);
                            acc[1][4 * j4 + 0] = fmaf(a1, wv.x, acc[1][4 * j4 + 0]);
                            acc[1][4 * j4 + 1] = fmaf(a1, wv.y, acc[1][4 * j4 + 1]);
                            acc[1][4 * j4 + 2] = fmaf(a1, wv.z, acc[1][4 * j4 + 2]);
                            acc[1][4 * j4 + 3] = fmaf(a1, wv.w, acc[1][4 * j4 + 3]);
                        }
                    }
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                uint4 pk[2], pb[2];
                __half2* hp = reinterpret_cast<__half2*>(pk);
                __half2* hb = reinterpret_cast<__half2*>(pb);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = g * 16 + 2 * j;
                    float a = acc[px][2 * j], b = acc[px][2 * j + 1];
                    a = a > 0.f ? a : a * sp[n];
                    b = b > 0.f ? b : b * sp[n + 1];
                    hp[j] = __floats2half2_rn(a, b);
                    const float2 yr = __half22float2(hp[j]);
                    hb[j] = __floats2half2_rn(fmaf(yr.x, ss[n], sbb[n]), fmaf(yr.y, ss[n + 1], sbb[n + 1]));
                }
                uint4* d0 = reinterpret_cast<uint4*>(y + o + px * 64 + g * 16);
                d0[0] = pk[0];
                d0[1] = pk[1];
                if (y_bn) {
                    uint4* d1 = reinterpret_cast<uint4*>(y_bn + o + px * 64 + g * 16);
                    d1[0] = pb[0];
                    d1[1] = pb[1];
                }
            }
        }
    }
}

// one thread per pixel (FR_STEM_PAIR=0; see the A/B note at the launch site in embedder.cu)
template <bool kU8>
__global__ void __launch_bounds__(128) arcface_stem_kernel(const void* __restrict__ in, int batch, const float* __restrict__ w /*[64][27]*/,
                                                           const float* __restrict__ bias, const float* __restrict__ prelu,
                                                           const float* __restrict__ bn_s, const float* __restrict__ bn_b,
                                                           __half* __restrict__ y, __half* __restrict__ y_bn) {
    constexpr int S = 112, Wp = S + 1, HpWp = Wp * Wp;
    __shared__ float4 ws[27][16];  // ws[k][n/4] = w[n..n+3][k]
    __shared__ float sb[64], sp[64], ss[64], sbb[64];
    for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
        const int k = i / 64, n = i % 64;
        reinterpret_cast<float*>(&ws[k][0])[n] = w[n * 27 + k];
    }
    if (threadIdx.x < 64) {
        sb[threadIdx.x] = bias[threadIdx.x];
        sp[threadIdx.x] = prelu[threadIdx.x];
        ss[threadIdx.x] = bn_s ? bn_s[threadIdx.x] : 1.f;
        sbb[threadIdx.x] = bn_b ? bn_b[threadIdx.x] : 0.f;
    }
    __syncthreads();
    // grid-stride over the pixels: the 6.9 KiB of weights staged above are amortised over several pixels per thread
    const long long total = static_cast<long long>(batch) * S * S;
    for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
         pix += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int img = static_cast<int>(pix / (S * S));
    const int rc = static_cast<int>(pix - static_cast<long long>(img) * S * S);
    const int r = rc / S, c = rc % S;
    float x[27];  // k = (ky*3 + kx)*3 + ch, ch in R,G,B order
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int rr = r + ky - 1, cc = c + kx - 1;
            const bool ok = rr >= 0 && rr < S && cc >= 0 && cc < S;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float v = 0.f;
                if (ok) {
                    if (kU8) {
                        const uint8_t u = static_cast<const uint8_t*>(in)[(static_cast<size_t>(img) * S * S + rr * S + cc) * 3 + (2 - ch)];
                        v = (static_cast<float>(u) - 127.5f) * 0.0078125f;
                    } else {
                        v = static_cast<const float*>(in)[(static_cast<size_t>(img) * 3 + ch) * S * S + rr * S + cc];
                    }
                }
                x[(ky * 3 + kx) * 3 + ch] = v;
            }
        }
    const size_t o = (static_cast<size_t>(img) * HpWp + r * Wp + c) * 64;
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = sb[g * 16 + j];
#pragma unroll
        for (int k = 0; k < 27; ++k) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 wv = ws[k][g * 4 + j4];
                acc[4 * j4 + 0] = fmaf(x[k], wv.x, acc[4 * j4 + 0]);
                acc[4 * j4 + 1] = fmaf(x[k], wv.y, acc[4 * j4 + 1]);
                acc[4 * j4 + 2] = fmaf(x[k], wv.z, acc[4 * j4 + 2]);
                acc[4 * j4 + 3] = fmaf(x[k], wv.w, acc[4 * j4 + 3]);
            }
        }
        uint4 pk[2], pb[2];
        __half2* hp = reinterpret_cast<__half2*>(pk);
        __half2* hb = reinterpret_cast<__half2*>(pb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = g * 16 + 2 * j;
            float a = acc[2 * j], b = acc[2 * j + 1];
            a = a > 0.f ? a : a * sp[n];
            b = b > 0.f ? b : b * sp[n + 1];
            hp[j] = __floats2half2_rn(a, b);
            const float2 yr = __half22float2(hp[j]);
            hb[j] = __floats2half2_rn(fmaf(yr.x, ss[n], sbb[n]), fmaf(yr.y, ss[n + 1], sbb[n + 1]));
        }
        uint4* d0 = reinterpret_cast<uint4*>(y + o + g * 16);
        d0[0] = pk[0];
        d0[1] = pk[1];
        if (y_bn) {
            uint4* d1 = reinterpret_cast<uint4*>(y_bn + o + g * 16);
            d1[0] = pb[0];
            d1[1] = pb[1];
        }
    }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// SEModule gate (model_irse.py:22-45): gate[img][c] = sigmoid(fc2 · relu(fc1 · mean_hw(u[img]))).
// u: [batch * HpWp, C] fp16 with zero pads, so the sum over all HpWp positions is the sum over the H*W pixels.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSeChunks = 16;  // position slices per image: the pooled sum is formed in two fixed-order stages (deterministic)
// stage 1: pool[(img * kSeChunks + chunk) * C + c] = sum of u over the chunk's positions. grid (batch, kSeChunks).
__global__ void __launch_bounds__(256) se_pool_kernel(const __half* __restrict__ u, int HpWp, int C, float* __restrict__ pool) {
    __shared__ float part[256 * 2];
    const int img = blockIdx.x, chunk = blockIdx.y;
    const int lanes = C / 2;             // threads covering one position (2 channels each)
    const int groups = 256 / lanes;      // position groups (C = 512 -> 1, C = 64 -> 8)
    const int cpair = threadIdx.x % lanes, grp = threadIdx.x / lanes;
    const int len = (HpWp + kSeChunks - 1) / kSeChunks;
    const int beg = chunk * len, end = min(HpWp, beg + len);
    float sx = 0.f, sy = 0.f;
    if (grp < groups) {
        const __half2* base = reinterpret_cast<const __half2*>(u + static_cast<size_t>(img) * HpWp * C) + cpair;
        for (int pos = beg + grp; pos < end; pos += groups) {
            const float2 v = __half22float2(base[static_cast<size_t>(pos) * lanes]);
            sx += v.x;
            sy += v.y;
        }
    }
    part[threadIdx.x * 2] = sx;
    part[threadIdx.x * 2 + 1] = sy;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int g = 0; g < groups; ++g) s += part[(g * lanes + (c >> 1)) * 2 + (c & 1)];
        pool[(static_cast<size_t>(img) * kSeChunks + chunk) * C + c] = s;
    }
}
// stage 2: one block per image
__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ pool, int HW, int C, const float* __restrict__ fc1,
                                                      const float* __restrict__ fc2, float* __restrict__ gate) {
    __shared__ float mean[512];
    __shared__ float hid[32];
    const int img = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < kSeChunks; ++k) s += pool[(static_cast<size_t>(img) * kSeChunks + k) * C + c];
        mean[c] = s / static_cast<float>(HW);
    }
    __syncthreads();
    const int hidden = C / 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < hidden; j += 8) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(fc1[j * C + c], mean[c], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) hid[j] = fmaxf(s, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < hidden; ++j) s = fmaf(fc2[c * hidden + j], hid[j], s);
        gate[img * C + c] = 1.f / (1.f + expf(-s));
    }
}

// y = u * gate + shortcut, plus the same side outputs as the conv epilogue (next unit's BN, subsampled copy).
// One thread per (position, 8 channels). res_mode: 1 = same geometry, 2 = subsample from (2H, 2W).
__global__ void __launch_bounds__(256) se_apply_kernel(const __half* __restrict__ u, const float* __restrict__ gate, int P, int H, int W, int C,
                                                       const __half* __restrict__ res, int res_mode, __half* __restrict__ y,
                                                       __half* __restrict__ y_bn, const float* __restrict__ bn_s, const float* __restrict__ bn_b,
                                                       __half* __restrict__ y_sub) {
    const int chunks = C / 8;
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(P) * chunks) return;
    const int p = static_cast<int>(t / chunks), n = static_cast<int>(t % chunks) * 8;
    const int Wp = W + 1, HpWp = (H + 1) * Wp;
    const int img = p / HpWp, rem = p - img * HpWp, r = rem / Wp, c = rem - r * Wp;
    if (r >= H || c >= W) return;
    size_t o_res = static_cast<size_t>(p);
    if (res_mode == 2) {
        const int W2p = 2 * W + 1, H2pW2p = (2 * H + 1) * W2p;
        o_res = static_cast<size_t>(img) * H2pW2p + (2 * r) * W2p + 2 * c;
    }
    const uint4 uv = *reinterpret_cast<const uint4*>(u + static_cast<size_t>(p) * C + n);
    const uint4 rv = *reinterpret_cast<const uint4*>(res + o_res * C + n);
    const __half2* uh = reinterpret_cast<const __half2*>(&uv);
    const __half2* rh = reinterpret_cast<const __half2*>(&rv);
    uint4 pk, pb;
    __half2* hp = reinterpret_cast<__half2*>(&pk);
    __half2* hb = reinterpret_cast<__half2*>(&pb);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 a = __half22float2(uh[j]), b = __half22float2(rh[j]);
        const float g0 = gate[img * C + n + 2 * j], g1 = gate[img * C + n + 2 * j + 1];
        hp[j] = __floats2half2_rn(fmaf(a.x, g0, b.x), fmaf(a.y, g1, b.y));
        if (y_bn) {
            const float2 yr = __half22float2(hp[j]);
            hb[j] = __floats2half2_rn(fmaf(yr.x, bn_s[n + 2 * j], bn_b[n + 2 * j]), fmaf(yr.y, bn_s[n + 2 * j + 1], bn_b[n + 2 * j + 1]));
        }
    }
    *reinterpret_cast<uint4*>(y + static_cast<size_t>(p) * C + n) = pk;
    if (y_bn) *reinterpret_cast<uint4*>(y_bn + static_cast<size_t>(p) * C + n) = pb;
    if (y_sub && !(r & 1) && !(c & 1)) {
        const int Wh = (W >> 1) + 1, HhWh = ((H >> 1) + 1) * Wh;
        *reinterpret_cast<uint4*>(y_sub + (static_cast<size_t>(img) * HhWh + (r >> 1) * Wh + (c >> 1)) * C + n) = pk;
    }
}

// output_layer tail: sum of the split-K partials of the folded Linear + bias, then F.normalize(p=2, dim=1, eps=1e-12)
// (model_irse.py:142-147,171). One block of 512 threads per image; fixed summation order (deterministic embeddings).
__global__ void __launch_bounds__(512) fc_reduce_l2norm_kernel(const float* __restrict__ partial, int splits, int batch,
                                                               const float* __restrict__ bias, float* __restrict__ out) {
    __shared__ float red[16];
    const int row = blockIdx.x, o = threadIdx.x;
    float v = bias[o];
    for (int s = 0; s < splits; ++s) v += partial[(static_cast<size_t>(s) * batch + row) * 512 + o];
    float sq = v * v;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) tot += red[i];
    out[static_cast<size_t>(row) * 512 + o] = v / fmaxf(sqrtf(tot), 1e-12f);
}

// debug/trace: shared-halo flat NHWC fp16 -> dense NCHW f32
__global__ void __launch_bounds__(256) unpack_nchw_kernel(const __half* __restrict__ src, int batch, int H, int W, int C, float* __restrict__ dst) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(batch) * C * H * W;
    if (t >= total) return;
    const int w = static_cast<int>(t % W), h = static_cast<int>((t / W) % H), c = static_cast<int>((t / (static_cast<long long>(W) * H)) % C);
    const int img = static_cast<int>(t / (static_cast<long long>(W) * H * C));
    const int Wp = W + 1, HpWp = (H + 1) * Wp;
    dst[t] = __half2float(src[(static_cast<size_t>(img) * HpWp + h * Wp + w) * C + c]);
}

}  // namespace frb
