// CUDA-core kernels around the tensor-core conv GEMM of the ArcFace IR-(SE)50 embedder
// (network spec: /root/reference conversion/arcface/model_irse.py; preprocessing: src/arcface.cpp:105-129).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx_sm100.cuh"

namespace frb {

// ---------------------------------------------------------------------------------------------------------------
// input_layer: Conv3x3(3->64, s1, p1) + BN (folded) + PReLU (model_irse.py:139-141), on the TENSOR cores: the 3x3x3 neighbourhood of a
// pixel is a K = 27 (padded to 32) row of an im2col operand that the CTA's 128 threads build directly in shared memory in the UMMA
// layout (K-major 128-byte rows, 128-byte swizzle: 16-byte chunk j of row r at chunk j ^ (r & 7)); one elected thread issues two
// tcgen05.mma (M = 128 pixels, N = 64 channels, K = 16 each) into a 64-column TMEM accumulator and the same threads run the epilogue
// (bias, PReLU, y and BN(y) stores). Input: either the tensor ArcFaceIR50::preprocessFaces produces (f32 planar R,G,B,
// (x-127.5)*0.0078125, src/arcface.cpp:118-125) or the u8 BGR HWC crop itself (same arithmetic on the fly); (u8 - 127.5) / 128 is exact
// in fp16; the folded weights are rounded to fp16 like every other layer's. Output (shared-halo flat NHWC, H = W = 112): y and
// y_bn = y * bn_s + bn_b (unit 0's pre-activation BN). The round-1 CUDA-core kernels were FMA-issue bound (1728 FMAs per pixel, 440-640 us
// at batch 256); this one is bound by its 256 B of output per pixel.
// Persistent: tile = 128 consecutive matrix rows of the output map, tile = blockIdx.x, + gridDim.x, ...; several CTAs per SM overlap
// one another's build / MMA / store phases.
// ---------------------------------------------------------------------------------------------------------------
template <bool kU8>
__global__ void __launch_bounds__(128) arcface_stem_tc_kernel(const void* __restrict__ in, int batch, const float* __restrict__ w /*[64][27]*/,
                                                              const float* __restrict__ bias, const float* __restrict__ prelu,
                                                              const float* __restrict__ bn_s, const float* __restrict__ bn_b,
                                                              __half* __restrict__ y, __half* __restrict__ y_bn) {
    constexpr int S = 112, Wp = S + 1, HpWp = Wp * Wp;
    __shared__ __align__(1024) uint8_t a_tile[128 * 128];
    __shared__ __align__(1024) uint8_t b_tile[64 * 128];
    __shared__ float sb[64], sp[64], ss[64], sbb[64];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    griddep_launch_dependents();
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<64>(&tmem_slot);
    for (int i = tid; i < 64 * 32; i += 128) {
        const int n = i >> 5, k = i & 31;
        const __half hv = __float2half_rn(k < 27 ? w[n * 27 + k] : 0.f);
        *reinterpret_cast<__half*>(b_tile + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = hv;
    }
    if (tid < 64) {
        sb[tid] = bias[tid];
        sp[tid] = prelu[tid];
        ss[tid] = bn_s ? bn_s[tid] : 1.f;
        sbb[tid] = bn_b ? bn_b[tid] : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    griddep_wait();  // the input (crops) and the output maps belong to the kernels before this one
    const uint32_t a_addr = smem_u32(a_tile), b_addr = smem_u32(b_tile);
    constexpr uint32_t idesc = umma_idesc(128, 64, 0, 0);
    const int P = batch * HpWp;
    const int tiles = (P + 127) / 128;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int p = tile * 128 + tid;
        const int img = p / HpWp, rem = p - img * HpWp;
        const int r = rem / Wp, c = rem - r * Wp;
        const bool valid = p < P && r < S && c < S;
        // ---- this pixel's im2col row: k = (ky*3 + kx)*3 + ch, ch in R,G,B order; 32 halves = 4 chunks
        __align__(16) __half row[32];
#pragma unroll
        for (int k = 27; k < 32; ++k) row[k] = __float2half_rn(0.f);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int rr = r + ky - 1, cc = c + kx - 1;
                const bool ok = valid && rr >= 0 && rr < S && cc >= 0 && cc < S;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float v = 0.f;
                    if (ok) {
                        if (kU8) {
                            const uint8_t u8v = static_cast<const uint8_t*>(in)[(static_cast<size_t>(img) * S * S + rr * S + cc) * 3 + (2 - ch)];
                            v = (static_cast<float>(u8v) - 127.5f) * 0.0078125f;
                        } else {
                            v = static_cast<const float*>(in)[(static_cast<size_t>(img) * 3 + ch) * S * S + rr * S + cc];
                        }
                    }
                    row[(ky * 3 + kx) * 3 + ch] = __float2half_rn(v);
                }
            }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(a_tile + tid * 128 + ((j ^ (tid & 7)) << 4)) = *reinterpret_cast<const uint4*>(row + j * 8);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            umma_f16_ss(tmem_base, umma_desc_sw128(a_addr), umma_desc_sw128(b_addr), idesc, 0u);
            umma_f16_ss(tmem_base, umma_desc_sw128(a_addr + 32), umma_desc_sw128(b_addr + 32), idesc, 1u);
            umma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        // Epilogue. A lane owns one pixel = one 128-byte output row; storing it from the lane directly would touch 32 different
        // 128-byte lines per store instruction (measured: the kernel was bound by exactly that, 440 us at batch 256). The rows are
        // staged through this warp's 32 rows of the (now free) operand tile instead and written back 4 whole rows per instruction.
        const int lane = tid & 31;
        uint8_t* stage = a_tile + warp * 32 * 128;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        uint4 pb[8];
#pragma unroll
        for (int cc = 0; cc < 64; cc += 16) {
            uint32_t raw[16];
            tmem_ld_32x32b_x16(taddr + cc, raw);
            tmem_ld_wait_x16(raw);
            uint4 pk[2];
            __half2* hp = reinterpret_cast<__half2*>(pk);
            __half2* hb = reinterpret_cast<__half2*>(&pb[cc / 8]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = cc + 2 * j;
                float a = __uint_as_float(raw[2 * j]) + sb[n], b = __uint_as_float(raw[2 * j + 1]) + sb[n + 1];
                a = a > 0.f ? a : a * sp[n];
                b = b > 0.f ? b : b * sp[n + 1];
                hp[j] = __floats2half2_rn(a, b);
                const float2 yr = __half22float2(hp[j]);
                hb[j] = __floats2half2_rn(fmaf(yr.x, ss[n], sbb[n]), fmaf(yr.y, ss[n + 1], sbb[n + 1]));
            }
            *reinterpret_cast<uint4*>(stage + lane * 128 + (((cc / 8) ^ (lane & 7)) << 4)) = pk[0];
            *reinterpret_cast<uint4*>(stage + lane * 128 + (((cc / 8 + 1) ^ (lane & 7)) << 4)) = pk[1];
        }
        const size_t o_warp = static_cast<size_t>(tile * 128 + warp * 32) * 64;  // halves; the warp's 32 rows are contiguous in y / y_bn
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int q = 4 * i + (lane >> 3);
            const uint4 v = *reinterpret_cast<const uint4*>(stage + q * 128 + (((lane & 7) ^ (q & 7)) << 4));
            if ((vmask >> q) & 1u) *reinterpret_cast<uint4*>(y + o_warp + q * 64 + (lane & 7) * 8) = v;
        }
        if (y_bn) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) = pb[j];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int q = 4 * i + (lane >> 3);
                const uint4 v = *reinterpret_cast<const uint4*>(stage + q * 128 + (((lane & 7) ^ (q & 7)) << 4));
                if ((vmask >> q) & 1u) *reinterpret_cast<uint4*>(y_bn + o_warp + q * 64 + (lane & 7) * 8) = v;
            }
        }
        tc_fence_before();
        __syncthreads();  // every TMEM read of this tile is done and the operand tile is free before the next one is built
    }
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// SEModule (model_irse.py:22-45): gate[img][c] = sigmoid(fc2 . relu(fc1 . mean_hw(u[img]))). The per-image channel sums come from the
// pooling partials the producing conv's epilogue wrote (pool[(group * 2 + seg) * C + c], group = matrix row / 32, see pool_store16 in
// conv_kernels.cuh): exact fixed-point sums, so embeddings are deterministic and independent of the batch position. One block per image.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) se_gate_kernel(const int* __restrict__ pool, int H, int W, int C, const float* __restrict__ fc1,
                                                      const float* __restrict__ fc2, float* __restrict__ gate) {
    __shared__ long long part[512];
    __shared__ float mean[512];
    __shared__ float hid[32];
    // The FC weights are static: fetch this thread's share BEFORE waiting for the producing conv, so their L2 round trips overlap the
    // pooling loads instead of following them (the kernel is a chain of dependent round trips: 7.8 us for a few thousand MACs).
    // C <= 512, hidden = C / 16 <= 32: warp w owns hidden units w and w + 16, thread c < C owns row c of fc2.
    const int hidden = C / 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float w1[2][16];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        const int j = warp + 16 * jj;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = lane + 32 * i;
            w1[jj][i] = (j < hidden && c < C) ? __ldg(fc1 + j * C + c) : 0.f;
        }
    }
    float4 w2[8];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4)
        w2[j4] = (threadIdx.x < C && j4 < hidden / 4) ? __ldg(reinterpret_cast<const float4*>(fc2 + static_cast<size_t>(threadIdx.x) * hidden) + j4)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    griddep_launch_dependents();
    griddep_wait();
    const int img = blockIdx.x;
    const int HpWp = (H + 1) * (W + 1);
    // ---- pooled mean: groups g_lo..g_hi overlap this image; of the first one only the part inside the image counts.
    //      512 threads = C channels x S group subsets (C * S == 512); exact integer sums, so the split does not change the result
    const int g_lo = (img * HpWp) >> 5, g_hi = ((img + 1) * HpWp - 1) >> 5;
    const int first_seg = (g_lo << 5) < img * HpWp ? 1 : 0;  // the group starts in the previous image: this image is its segment 1
    const int S = 512 / C;
    {
        const int c = threadIdx.x % C, sub = threadIdx.x / C;
        long long s = 0;
        int g = g_lo + sub;
        if (g == g_lo && g <= g_hi) {
            s += pool[(static_cast<size_t>(g) * 2 + first_seg) * C + c];
            g += S;
        }
        for (; g + 3 * S <= g_hi; g += 4 * S) {  // four loads in flight
            const int a0 = pool[static_cast<size_t>(g) * 2 * C + c], a1 = pool[static_cast<size_t>(g + S) * 2 * C + c];
            const int a2 = pool[static_cast<size_t>(g + 2 * S) * 2 * C + c], a3 = pool[static_cast<size_t>(g + 3 * S) * 2 * C + c];
            s += static_cast<long long>(a0) + a1 + a2 + a3;
        }
        for (; g <= g_hi; g += S) s += pool[static_cast<size_t>(g) * 2 * C + c];
        part[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x < C) {
        long long s = 0;
        for (int sub = 0; sub < S; ++sub) s += part[sub * C + threadIdx.x];
        mean[threadIdx.x] = static_cast<float>(static_cast<double>(s) * (1.0 / 16384.0) / static_cast<double>(H * W));  // kPoolScale
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        const int j = warp + 16 * jj;
        if (j < hidden) {  // warp-uniform
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {  // same order as c = lane, lane + 32, ...
                const int c = lane + 32 * i;
                if (c < C) s = fmaf(w1[jj][i], mean[c], s);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) hid[j] = fmaxf(s, 0.f);
        }
    }
    __syncthreads();
    if (threadIdx.x < C) {
        float s = 0.f;  // hidden is a multiple of 4
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            if (j4 >= hidden / 4) break;
            const float4 w4 = w2[j4];
            s = fmaf(w4.x, hid[4 * j4], s);
            s = fmaf(w4.y, hid[4 * j4 + 1], s);
            s = fmaf(w4.z, hid[4 * j4 + 2], s);
            s = fmaf(w4.w, hid[4 * j4 + 3], s);
        }
        gate[static_cast<size_t>(img) * C + threadIdx.x] = 1.f / (1.f + expf(-s));
    }
}

// y = u * gate + shortcut (model_irse.py:62-66), plus the side outputs of the conv epilogue (next unit's BatchNorm copy, subsampled
// copy). Pure streaming: one item = (position, 8 channels). The grid stride is a multiple of 256 and C / 8 divides 256, so a thread
// keeps the same channel chunk for all its items: its BatchNorm scale / shift live in registers; two items in flight per thread.
// res_mode: 1 = same geometry, 2 = subsample from (2H, 2W).
__global__ void __launch_bounds__(256, 3) se_apply_kernel(const __half* __restrict__ u, const float* __restrict__ gate, int P, int H, int W, int C,
                                                       const __half* __restrict__ res, int res_mode, __half* __restrict__ y,
                                                       __half* __restrict__ y_bn, const float* __restrict__ bn_s, const float* __restrict__ bn_b,
                                                       __half* __restrict__ y_sub) {
    griddep_launch_dependents();
    const int chunks = C / 8;
    const int Wp = W + 1, HpWp = (H + 1) * Wp;
    const int Wh = (W >> 1) + 1, HhWh = ((H >> 1) + 1) * Wh;
    const int n = (threadIdx.x % chunks) * 8;            // this thread's channels, fixed
    const int pos_per_pass = (gridDim.x * 256) / chunks;  // positions the whole grid covers per pass
    float sc[8], bi[8];
    if (y_bn) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(bn_s + n)), s1 = __ldg(reinterpret_cast<const float4*>(bn_s + n) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bn_b + n)), b1 = __ldg(reinterpret_cast<const float4*>(bn_b + n) + 1);
        sc[0] = s0.x, sc[1] = s0.y, sc[2] = s0.z, sc[3] = s0.w, sc[4] = s1.x, sc[5] = s1.y, sc[6] = s1.z, sc[7] = s1.w;
        bi[0] = b0.x, bi[1] = b0.y, bi[2] = b0.z, bi[3] = b0.w, bi[4] = b1.x, bi[5] = b1.y, bi[6] = b1.z, bi[7] = b1.w;
    }
    griddep_wait();  // the BatchNorm parameters above are static; u, gate and the residual are not
    for (int p0 = (blockIdx.x * 256 + threadIdx.x) / chunks; p0 < P; p0 += 2 * pos_per_pass) {
        uint4 uv[2], rv[2];
        float4 g0[2], g1[2];
        int pp[2], im[2], rr[2], cc[2];
        bool ok[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            pp[k] = p0 + k * pos_per_pass;
            im[k] = pp[k] / HpWp;
            const int rem = pp[k] - im[k] * HpWp;
            rr[k] = rem / Wp;
            cc[k] = rem - rr[k] * Wp;
            ok[k] = pp[k] < P && rr[k] < H && cc[k] < W;
            if (ok[k]) {
                size_t o_res = static_cast<size_t>(pp[k]);
                if (res_mode == 2) {
                    const int W2p = 2 * W + 1, H2pW2p = (2 * H + 1) * W2p;
                    o_res = static_cast<size_t>(im[k]) * H2pW2p + (2 * rr[k]) * W2p + 2 * cc[k];
                }
                uv[k] = __ldg(reinterpret_cast<const uint4*>(u + static_cast<size_t>(pp[k]) * C + n));
                rv[k] = __ldg(reinterpret_cast<const uint4*>(res + o_res * C + n));
                g0[k] = __ldg(reinterpret_cast<const float4*>(gate + static_cast<size_t>(im[k]) * C + n));
                g1[k] = __ldg(reinterpret_cast<const float4*>(gate + static_cast<size_t>(im[k]) * C + n) + 1);
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!ok[k]) continue;
            const float g[8] = {g0[k].x, g0[k].y, g0[k].z, g0[k].w, g1[k].x, g1[k].y, g1[k].z, g1[k].w};
            const __half2* uh = reinterpret_cast<const __half2*>(&uv[k]);
            const __half2* rh = reinterpret_cast<const __half2*>(&rv[k]);
            uint4 pk, pb;
            __half2* hp = reinterpret_cast<__half2*>(&pk);
            __half2* hb = reinterpret_cast<__half2*>(&pb);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 a = __half22float2(uh[j]), b = __half22float2(rh[j]);
                hp[j] = __floats2half2_rn(fmaf(a.x, g[2 * j], b.x), fmaf(a.y, g[2 * j + 1], b.y));
            }
            const size_t o = static_cast<size_t>(pp[k]) * C + n;
            *reinterpret_cast<uint4*>(y + o) = pk;
            if (y_bn) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 yr = __half22float2(hp[j]);
                    hb[j] = __floats2half2_rn(fmaf(yr.x, sc[2 * j], bi[2 * j]), fmaf(yr.y, sc[2 * j + 1], bi[2 * j + 1]));
                }
                *reinterpret_cast<uint4*>(y_bn + o) = pb;
            }
            if (y_sub && !(rr[k] & 1) && !(cc[k] & 1))
                *reinterpret_cast<uint4*>(y_sub + (static_cast<size_t>(im[k]) * HhWh + (rr[k] >> 1) * Wh + (cc[k] >> 1)) * C + n) = pk;
        }
    }
}

// output_layer tail: sum of the split-K partials of the folded Linear + bias, then F.normalize(p=2, dim=1, eps=1e-12)
// (model_irse.py:142-147,171). One block of 512 threads per image; fixed summation order (deterministic embeddings).
__global__ void __launch_bounds__(512) fc_reduce_l2norm_kernel(const float* __restrict__ partial, int splits, int batch,
                                                               const float* __restrict__ bias, float* __restrict__ out) {
    __shared__ float red[16];
    griddep_launch_dependents();
    griddep_wait();
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
