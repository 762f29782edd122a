// Device code of the cosine-similarity search:
//   cosine_topk_coarse<CG>  fused  Q x 512 · (N x 512)^T  on tcgen05 tensor cores (fp16 in, fp32 accumulate in TMEM)
//                           + running top-KC per query in registers; the similarity matrix is never written.
//   topk_rerank_kernel      merges the per-CTA candidates, re-scores the KC survivors per query in exact fp32
//                           from the fp32 master rows and orders them by (score desc, row asc).
//   sims_kernel             exact fp32 dense similarities (MatMul::calculate, /root/reference src/matmul.cpp:36-77).
//   topk_dense_kernel       top-k of a dense similarity matrix (ArcFaceIR50::getOutputs, src/arcface.cpp:203-217).
//   topk_merge_kernel       merge of per-shard (score, idx) lists after the cross-GPU all-gather.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx_sm100.cuh"

namespace frb {

constexpr int kDim = 512;          // rec_outputDim, app/config.json:16
constexpr int kKBlocks = 8;        // 512 / 64 : one 128-byte swizzle span of fp16 per k-block
constexpr int kTileRows = 256;     // gallery rows per accumulator tile (UMMA N)
constexpr int kQRows = 128;        // queries per CTA (UMMA M per CTA = TMEM lanes)
constexpr int kKC = 8;             // coarse candidates kept per query per CTA (and re-scored per query)
constexpr int kQTileBytes = kQRows * 128;            // one k-block of the query operand: 16 KiB
constexpr int kQBytes = kKBlocks * kQTileBytes;      // 128 KiB
constexpr int kHalfTileBytes = 128 * 128;            // 128 gallery rows x 64 fp16 : one TMA box, 16 KiB
constexpr int kSearchThreads = 256;                  // warps 0-3: TMA / MMA / TMEM alloc / idle, warps 4-7: epilogue
constexpr int kRingBytes = 96 * 1024;                // gallery stage ring

template <int CG>
struct CoarseCfg {
    static constexpr int kStageBytes = (kTileRows / CG) * 128;   // bytes this CTA loads per k-block
    static constexpr int kStages = kRingBytes / kStageBytes;     // 3 (single CTA) or 6 (CTA pair)
    static constexpr int kSmemBytes = 1024 /*align slack*/ + kQBytes + kRingBytes + 256 /*barriers*/;
};

__device__ __forceinline__ bool better(float sa, int64_t ia, float sb, int64_t ib) {
    return sa > sb || (sa == sb && ia < ib);
}

// insert (v, id) into a descending list of kKC entries held in registers; ties keep the earlier entry first
__device__ __forceinline__ void topk_insert(float (&s)[kKC], int (&ix)[kKC], float v, int id) {
    s[kKC - 1] = v;
    ix[kKC - 1] = id;
#pragma unroll
    for (int t = kKC - 1; t > 0; --t) {
        if (s[t] > s[t - 1]) {
            float ts = s[t];
            s[t] = s[t - 1];
            s[t - 1] = ts;
            int ti = ix[t];
            ix[t] = ix[t - 1];
            ix[t - 1] = ti;
        }
    }
}

// ----------------------------------------------------------------------------------------------------------
// Fused coarse search. Grid: CG * units CTAs (cluster of CG). Unit u scans gallery tiles u, u+units, ...
// q: nq x 512 f32 (device). CTA rank r of a pair owns queries [128 r, 128 r + 128).
// cand_s / cand_i: [units][CG*128][kKC]  (score, local row) per unit and query, descending, (-inf,-1) padded.
// ----------------------------------------------------------------------------------------------------------
template <int CG>
__global__ void __launch_bounds__(kSearchThreads, 1)
cosine_topk_coarse(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ q, int nq, long long n_rows, int num_tiles,
                   float* __restrict__ cand_s, int* __restrict__ cand_i) {
    using Cfg = CoarseCfg<CG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* q_smem = smem;
    uint8_t* ring = smem + kQBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kRingBytes);  // [kStages]  (used in the leader CTA)
    uint64_t* empty_bar = full_bar + Cfg::kStages;                        // [kStages]
    uint64_t* tfull_bar = empty_bar + Cfg::kStages;                       // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;                                 // [2] accumulator drained (leader CTA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int unit = blockIdx.x / CG;
    const int num_units = gridDim.x / CG;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], CG * 4);  // one arrive per epilogue warp of every CTA of the unit
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if (CG == 2) tmem_alloc_pair<512>(tmem_slot);
        else tmem_alloc<512>(tmem_slot);
    }

    // ---- stage this CTA's 128 queries: f32 global -> fp16, K-major, 128-byte swizzled (the layout TMA would write)
    {
        const int q_first = static_cast<int>(cta_rank) * kQRows;
        for (int g = threadIdx.x; g < kQRows * 64; g += kSearchThreads) {
            const int r = g >> 6;       // query row inside the CTA
            const int ch = g & 63;      // 16-byte (8 x fp16) chunk along K
            const int kb = ch >> 3, c = ch & 7;
            uint4 packed = make_uint4(0u, 0u, 0u, 0u);
            if (q_first + r < nq) {
                const float4* src = reinterpret_cast<const float4*>(q + static_cast<size_t>(q_first + r) * kDim + ch * 8);
                const float4 a = __ldg(src), b = __ldg(src + 1);
                __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
                __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
                packed.x = *reinterpret_cast<uint32_t*>(&h0);
                packed.y = *reinterpret_cast<uint32_t*>(&h1);
                packed.z = *reinterpret_cast<uint32_t*>(&h2);
                packed.w = *reinterpret_cast<uint32_t*>(&h3);
            }
            *reinterpret_cast<uint4*>(q_smem + kb * kQTileBytes + r * 128 + ((c ^ (r & 7)) << 4)) = packed;
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            const uint32_t leader_full0 = (CG == 2) ? mapa_u32(smem_u32(&full_bar[0]), 0) : 0u;
            for (int t = unit; t < num_tiles; t += num_units) {
                const int row0 = t * kTileRows;
                for (int kb = 0; kb < kKBlocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* dst = ring + stage * Cfg::kStageBytes;
                    if (CG == 1) {
                        mbar_expect_tx(&full_bar[stage], 2 * kHalfTileBytes);
                        tma_load_2d(dst, &tmap, &full_bar[stage], kb * 64, row0, kEvictFirst);
                        tma_load_2d(dst + kHalfTileBytes, &tmap, &full_bar[stage], kb * 64, row0 + 128, kEvictFirst);
                    } else {
                        // the leader's barrier collects the bytes of both CTAs' halves
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kHalfTileBytes);
                        tma_load_2d_pair(dst, &tmap, leader_full0 + stage * 8, kb * 64, row0 + static_cast<int>(cta_rank) * 128,
                                         kEvictFirst);
                    }
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, one thread) =====================
        if (cta_rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc(kQRows * CG, kTileRows, 0, 0);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int t = unit; t < num_tiles; t += num_units, ++it) {
                const int buf = it & 1;
                mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * kTileRows;
                for (int kb = 0; kb < kKBlocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(q_smem + kb * kQTileBytes);
                    const uint32_t b_addr = smem_u32(ring + stage * Cfg::kStageBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 4 x (K = 16) inside the 128-byte swizzle span
                        const uint64_t da = umma_desc_sw128(a_addr + k * 32);
                        const uint64_t db = umma_desc_sw128(b_addr + k * 32);
                        if (CG == 2) umma_f16_ss_pair(d_tmem, da, db, idesc, (kb | k) != 0);
                        else umma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
                    }
                    if (CG == 2) umma_commit_pair(&empty_bar[stage], 0x3);  // frees the slot in both CTAs
                    else umma_commit(&empty_bar[stage]);
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (CG == 2) umma_commit_pair(&tfull_bar[buf], 0x3);
                else umma_commit(&tfull_bar[buf]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: running top-KC per query, in registers =====================
        float best_s[kKC];
        int best_i[kKC];
#pragma unroll
        for (int j = 0; j < kKC; ++j) {
            best_s[j] = -INFINITY;
            best_i[j] = -1;
        }
        const int ew = warp & 3;  // TMEM lane quarter this warp may access
        const uint32_t tempty_leader0 = (CG == 2) ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;
        int it = 0;
        for (int t = unit; t < num_tiles; t += num_units, ++it) {
            const int buf = it & 1;
            mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * kTileRows;
            const long long row0 = static_cast<long long>(t) * kTileRows;
            const int valid = (n_rows - row0 >= kTileRows) ? kTileRows : static_cast<int>(n_rows - row0);
#pragma unroll 1
            for (int c = 0; c < kTileRows / 32; ++c) {
                uint32_t raw[32];
                tmem_ld_32x32b_x32(taddr + c * 32, raw);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                if (valid < kTileRows) {  // last, partial tile: rows beyond the gallery are TMA zero fill
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j >= valid) v[j] = -INFINITY;
                }
                float m = v[0];
#pragma unroll
                for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
                if (m > best_s[kKC - 1]) {
                    const int base = static_cast<int>(row0) + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (v[j] > best_s[kKC - 1]) topk_insert(best_s, best_i, v[j], base + j);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(tempty_leader0 + buf * 8);
                else mbar_arrive(&tempty_bar[buf]);
            }
        }
        const int qrow = static_cast<int>(cta_rank) * kQRows + ew * 32 + lane;
        const size_t o = (static_cast<size_t>(unit) * (CG * kQRows) + qrow) * kKC;
#pragma unroll
        for (int j = 0; j < kKC; ++j) {
            cand_s[o + j] = best_s[j];
            cand_i[o + j] = best_i[j];
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        if (CG == 2) tmem_dealloc_pair<512>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

// ----------------------------------------------------------------------------------------------------------
// exact fp32 dot of two 512-vectors by one warp, fixed summation order (shared by every exact-score producer
// so that sims, re-rank and dense top-k agree bit for bit).  a: 16 registers per lane, element (lane + 32 i) * 4 + e.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load512(const float* __restrict__ p, int lane, float4 (&r)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = __ldg(reinterpret_cast<const float4*>(p) + lane + 32 * i);
}
__device__ __forceinline__ float dot512(const float4 (&a)[4], const float4 (&b)[4]) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc = fmaf(a[i].x, b[i].x, acc);
        acc = fmaf(a[i].y, b[i].y, acc);
        acc = fmaf(a[i].z, b[i].z, acc);
        acc = fmaf(a[i].w, b[i].w, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// block-wide selection of the K best of `count` (score, idx) entries in shared memory, order (score desc, idx asc).
// Entries with idx < 0 are ignored. Selected entries are consumed (idx set to -1). Result in out_s/out_i (shared, K entries,
// (-inf,-1) padded). All threads of the block must call; blockDim.x multiple of 32, <= 1024.
__device__ inline void block_select(float* cs, long long* ci, int count, int K, float* out_s, long long* out_i, float* red_s,
                                    long long* red_i, int* red_p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int round = 0; round < K; ++round) {
        float bs = -INFINITY;
        long long bi = -1;
        int bp = -1;
        for (int p = threadIdx.x; p < count; p += blockDim.x) {
            const long long id = ci[p];
            if (id < 0) continue;
            const float s = cs[p];
            if (bp < 0 || better(s, id, bs, bi)) {
                bs = s;
                bi = id;
                bp = p;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (op >= 0 && (bp < 0 || better(os, oi, bs, bi))) {
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
            for (int w = 1; w < nwarps; ++w) {
                if (red_p[w] >= 0 && (bp < 0 || better(red_s[w], red_i[w], bs, bi))) {
                    bs = red_s[w];
                    bi = red_i[w];
                    bp = red_p[w];
                }
            }
            if (bp >= 0) {
                out_s[round] = bs;
                out_i[round] = bi;
                ci[bp] = -1;
            } else {
                out_s[round] = -INFINITY;
                out_i[round] = -1;
            }
        }
        __syncthreads();
    }
}

constexpr int kSelThreads = 256;
constexpr int kSelMaxCand = 148 * kKC;  // one candidate list per CTA of the coarse kernel at most

// One block per query: merge `units` coarse lists, exact fp32 re-score of the kKC best, final order, write top-k.
// out_s: nq x k, out_i: nq x k (row_offset + local row), padded with (-inf, -1).
__global__ void __launch_bounds__(kSelThreads) topk_rerank_kernel(const float* __restrict__ cand_s, const int* __restrict__ cand_i,
                                                                  int units, int q_stride, const float* __restrict__ q,
                                                                  const float* __restrict__ rows, int k, long long row_offset,
                                                                  float* __restrict__ out_s, long long* __restrict__ out_i) {
    __shared__ float cs[kSelMaxCand];
    __shared__ long long ci[kSelMaxCand];
    __shared__ float sel_s[kKC];
    __shared__ long long sel_i[kKC];
    __shared__ float red_s[32];
    __shared__ long long red_i[32];
    __shared__ int red_p[32];
    const int qi = blockIdx.x;
    const int count = units * kKC;
    for (int p = threadIdx.x; p < count; p += blockDim.x) {
        const int u = p / kKC, j = p % kKC;
        const size_t o = (static_cast<size_t>(u) * q_stride + qi) * kKC + j;
        cs[p] = cand_s[o];
        ci[p] = cand_i[o];
    }
    __syncthreads();
    block_select(cs, ci, count, kKC, sel_s, sel_i, red_s, red_i, red_p);
    // exact scores: warp w re-scores candidate w
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < kKC) {
        const long long id = sel_i[warp];
        float s = -INFINITY;
        if (id >= 0) {
            float4 a[4], b[4];
            load512(q + static_cast<size_t>(qi) * kDim, lane, a);
            load512(rows + static_cast<size_t>(id) * kDim, lane, b);
            s = dot512(a, b);
        }
        if (lane == 0) sel_s[warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // insertion sort of kKC entries by (exact score desc, row asc); invalid entries last
        for (int a = 1; a < kKC; ++a) {
            const float s = sel_s[a];
            const long long id = sel_i[a];
            int b = a - 1;
            while (b >= 0 && id >= 0 && (sel_i[b] < 0 || better(s, id, sel_s[b], sel_i[b]))) {
                sel_s[b + 1] = sel_s[b];
                sel_i[b + 1] = sel_i[b];
                --b;
            }
            sel_s[b + 1] = s;
            sel_i[b + 1] = id;
        }
        for (int j = 0; j < k; ++j) {
            const bool ok = j < kKC && sel_i[j] >= 0;
            out_s[static_cast<size_t>(qi) * k + j] = ok ? sel_s[j] : -INFINITY;
            out_i[static_cast<size_t>(qi) * k + j] = ok ? sel_i[j] + row_offset : -1;
        }
    }
}

// exact dense similarities: out[i * n + j] = <q_i, row_j>. One warp per gallery row, queries staged in shared memory
// in chunks of kSimsQ. grid.x over rows (grid-stride), grid.y over query chunks.
constexpr int kSimsQ = 16;
constexpr int kSimsThreads = 256;
__global__ void __launch_bounds__(kSimsThreads) sims_kernel(const float* __restrict__ rows, long long n, const float* __restrict__ q,
                                                           int nq, float* __restrict__ out) {
    __shared__ float4 qs[kSimsQ][128];
    const int q0 = blockIdx.y * kSimsQ;
    const int qn = min(kSimsQ, nq - q0);
    for (int g = threadIdx.x; g < qn * 128; g += blockDim.x)
        qs[g >> 7][g & 127] = __ldg(reinterpret_cast<const float4*>(q + static_cast<size_t>(q0) * kDim) + g);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    for (long long j = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n; j += warps) {
        float4 b[4];
        load512(rows + static_cast<size_t>(j) * kDim, lane, b);
        for (int i = 0; i < qn; ++i) {
            float4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = qs[i][lane + 32 * u];
            const float s = dot512(a, b);
            if (lane == 0) out[static_cast<size_t>(q0 + i) * n + j] = s;
        }
    }
}

// top-k of each row of a dense similarity matrix (first maximum wins on ties = std::max_element, src/arcface.cpp:210)
__global__ void __launch_bounds__(kSelThreads) topk_dense_kernel(const float* __restrict__ sims, long long n, int k,
                                                                 long long row_offset, float* __restrict__ out_s,
                                                                 long long* __restrict__ out_i) {
    __shared__ float cs[kSelThreads * kKC];
    __shared__ long long ci[kSelThreads * kKC];
    __shared__ float sel_s[kKC];
    __shared__ long long sel_i[kKC];
    __shared__ float red_s[32];
    __shared__ long long red_i[32];
    __shared__ int red_p[32];
    const int qi = blockIdx.x;
    float bs[kKC];
    int bi[kKC];
#pragma unroll
    for (int j = 0; j < kKC; ++j) {
        bs[j] = -INFINITY;
        bi[j] = -1;
    }
    const float* row = sims + static_cast<size_t>(qi) * n;
    for (long long j = threadIdx.x; j < n; j += blockDim.x) {
        const float v = row[j];
        if (bi[kKC - 1] < 0 || v > bs[kKC - 1]) {
            if (!(v != v)) topk_insert(bs, bi, v, static_cast<int>(j));
        }
    }
#pragma unroll
    for (int j = 0; j < kKC; ++j) {
        cs[threadIdx.x * kKC + j] = bs[j];
        ci[threadIdx.x * kKC + j] = bi[j];
    }
    __syncthreads();
    block_select(cs, ci, kSelThreads * kKC, k, sel_s, sel_i, red_s, red_i, red_p);
    if (threadIdx.x < k) {
        const long long id = sel_i[threadIdx.x];
        out_s[static_cast<size_t>(qi) * k + threadIdx.x] = sel_s[threadIdx.x];
        out_i[static_cast<size_t>(qi) * k + threadIdx.x] = id >= 0 ? id + row_offset : -1;
    }
}

// merge of n_parts per-shard results (each nq x k, global indices) -> nq x k, order (score desc, idx asc)
__global__ void __launch_bounds__(64) topk_merge_kernel(const float* __restrict__ ps, const long long* __restrict__ pi, int n_parts,
                                                        int nq, int k, float* __restrict__ out_s, long long* __restrict__ out_i) {
    __shared__ float cs[64 * kKC];
    __shared__ long long ci[64 * kKC];
    __shared__ float sel_s[kKC];
    __shared__ long long sel_i[kKC];
    __shared__ float red_s[32];
    __shared__ long long red_i[32];
    __shared__ int red_p[32];
    const int qi = blockIdx.x;
    const int count = n_parts * k;
    for (int p = threadIdx.x; p < count; p += blockDim.x) {
        const int part = p / k, j = p % k;
        const size_t o = (static_cast<size_t>(part) * nq + qi) * k + j;
        cs[p] = ps[o];
        ci[p] = pi[o];
    }
    __syncthreads();
    block_select(cs, ci, count, k, sel_s, sel_i, red_s, red_i, red_p);
    if (threadIdx.x < k) {
        out_s[static_cast<size_t>(qi) * k + threadIdx.x] = sel_s[threadIdx.x];
        out_i[static_cast<size_t>(qi) * k + threadIdx.x] = sel_i[threadIdx.x];
    }
}

// ---------------------------------------------------------------- synthetic gallery rows (bench / tests)
// Counter-based and integer-exact so that oracle/search_oracle.py (synth_rows) regenerates any row bit for bit:
// element c of global row g: four 16-bit fields of mix64(mix64(seed ^ g * C1) + c) summed, centred (Irwin-Hall ~ normal),
// then the row is divided by its exact integer L2 norm in double precision.
__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline int synth_elem(uint64_t row_key, int c) {
    const uint64_t h = mix64(row_key + static_cast<uint64_t>(c));
    return static_cast<int>((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48)) - 131070;
}
__global__ void __launch_bounds__(256) synth_rows_kernel(float* __restrict__ rows32, __half* __restrict__ rows16, long long n,
                                                         unsigned long long seed, long long row_offset) {
    const int lane = threadIdx.x & 31;
    const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    for (long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps) {
        const uint64_t key = mix64(seed ^ (static_cast<uint64_t>(r + row_offset) * 0xD6E8FEB86659FD93ull));
        int v[16];
        long long ss = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            v[i] = synth_elem(key, lane + 32 * i);
            ss += static_cast<long long>(v[i]) * v[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const double scale = ss > 0 ? 1.0 / sqrt(static_cast<double>(ss)) : 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float f = static_cast<float>(static_cast<double>(v[i]) * scale);
            const size_t o = static_cast<size_t>(r) * kDim + lane + 32 * i;
            rows32[o] = f;
            rows16[o] = __float2half_rn(f);
        }
    }
}

__global__ void __launch_bounds__(256) f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n4) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a);
        o.y = *reinterpret_cast<uint32_t*>(&b);
        reinterpret_cast<uint2*>(dst)[i] = o;
    }
}

}  // namespace frb
