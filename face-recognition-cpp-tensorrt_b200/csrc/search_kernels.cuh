// Device code of the cosine-similarity search:
//   prep_queries_kernel<F8>      the scan's query operand (fp16 or scaled e4m3, row-major) + the per-query margin, once per search.
//   cosine_topk_coarse<CG,KSEL,F8,APP>  fused  Q x 512 · (N x 512)^T  on tcgen05 tensor cores (fp16 or e4m3 in, fp32 accumulate in
//                                TMEM) + a running candidate set per query; the similarity matrix is never written.
//                                APP (top-1): predicated appends to per-thread buffers behind a shared per-query threshold;
//                                otherwise (k > 1): sorted candidate lists in registers.
//   append_rerank_kernel / topk_rerank_kernel   keep every candidate whose coarse score is within the margin of the (k-th) best,
//                                re-score those in exact fp32 from the fp32 master rows and order them by (score desc, row asc).
//                                Queries whose candidate set could be incomplete are flagged and recomputed by the exact scan.
//   exact_scan_kernel / exact_merge_kernel   exact fp32 scan (small galleries, flagged queries, FR_PATH_EXACT).
//   sims_kernel                  exact fp32 dense similarities (MatMul::calculate, /root/reference src/matmul.cpp:36-77).
//   topk_merge_kernel            merge of per-shard (score, idx) lists after the cross-GPU all-gather.
//   make_scan_copy_kernel / make_f8_copy_kernel   the resident scan copies and the norm bounds the margins scale with.
// Ordering everywhere: (score descending, row index ascending) — for k = 1 this is std::max_element's "first maximum"
// (ArcFaceIR50::getOutputs, /root/reference src/arcface.cpp:210).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx_sm100.cuh"

namespace frb {

constexpr int kDim = 512;          // rec_outputDim, app/config.json:16
constexpr int kKBlocks = 8;        // fp16 scan: 512 / 64 : one 128-byte swizzle span per k-block (fp8 scan: 512 / 128 = 4)
constexpr int kTileRows = 256;     // gallery rows per accumulator tile (UMMA N)
constexpr int kQRows = 128;        // queries per CTA (UMMA M per CTA = TMEM lanes)
constexpr int kTopkMax = 8;        // FR_TOPK_MAX
constexpr int kQTileBytes = kQRows * 128;            // one k-block of the query operand: 16 KiB
constexpr int kQBytes = kKBlocks * kQTileBytes;      // 128 KiB
constexpr int kHalfTileBytes = 128 * 128;            // 128 gallery rows x 64 fp16 : one TMA box, 16 KiB
constexpr int kEpiWarps = 8;                         // 2 column halves x 4 TMEM lane quarters
constexpr int kSearchThreads = 128 + kEpiWarps * 32; // warps 0-3: TMA / MMA / TMEM alloc / idle, warps 4-11: epilogue
// The append epilogue of the e4m3 pair scan runs SIXTEEN epilogue warps (4 column quarters x 4 TMEM lane quarters): under the wide
// certified margin its eight warps were busy 77 % of the time at 29 % issue-slot use (two warps per scheduler, every append a chain of
// dependent votes and branches): latency-bound, not throughput-bound (profiles/r02_f8_unknown_1250k_before_ncu_source.txt), while an
// e4m3 tile leaves the epilogue half the time an fp16 tile does. 640 threads x 96 registers still fit one CTA per SM.
#ifndef FR_AB_EW  // A/B switches (face-recognition-cpp-tensorrt_b200/build.py, tools/ab_search.py)
#define FR_AB_EW 16
#endif
#ifndef FR_AB_DEFER
#define FR_AB_DEFER 2
#endif
#ifndef FR_AB_GBLAG
#define FR_AB_GBLAG 1
#endif
template <int CG, bool F8, bool APP>
constexpr int coarse_epi_warps() {
    return (CG == 2 && F8 && APP) ? FR_AB_EW : kEpiWarps;
}
template <int CG, bool F8, bool APP>
constexpr int coarse_threads() {
    return 128 + coarse_epi_warps<CG, F8, APP>() * 32;
}
constexpr int kRingBytes = 96 * 1024;                // gallery stage ring
constexpr int kLdCols = 16;                          // accumulator columns per tcgen05.ld
// |coarse - exact| <= kCoarseEps * |q| * |row|: fp16 round-to-nearest of both operands (2 * 2^-11) plus the tensor core's fp32
// accumulation, bounded through Cauchy-Schwarz. Candidates within 2 eps of the k-th best coarse score are re-scored exactly.
constexpr float kCoarseEps = 1.25e-3f;
// ... which holds while operands stay in fp16's normal range. Subnormal components carry an ABSOLUTE error of 2^-25 each, i.e. up to
// sqrt(512) 2^-25 |other operand| per score; that stays inside kCoarseEps |q| |g| while the norm is >= sqrt(512) 2^-25 / kCoarseEps =
// 5.4e-4. Galleries whose largest row norm is below kF16MinNorm (or with a component beyond 65504) are searched by the exact fp32 scan,
// such queries are recomputed by it.
constexpr float kF16MinNorm = 1e-3f;

// fp8 (e4m3) scan copy: rows and queries are multiplied by kF8Scale before the conversion (unit-norm components ~0.044 land in
// e4m3's normal range), accumulators are kF8Scale^2 x the cosine. Round-to-nearest e4m3 has no useful error bound (2^-4 per operand
// through Cauchy-Schwarz, and structured rows make the "independent errors" reading false), so both operands are rounded
// STOCHASTICALLY (f8_round_dither: to one of the two neighbouring e4m3 values, up with probability = the fractional position,
// uniform variates from a counter-based hash of (seed, row, column)). Then, for ANY fixed query q and row g,
//     coarse - exact = sum_i qhat_i r_i(g) + sum_i g_i r_i(q)     (r = rounding error, independent, zero-mean, |range| = the e4m3 step u_i)
// and Hoeffding's inequality gives  P(exact - coarse >= t) <= exp(-2 t^2 / V),
//     V = sum qbar_i^2 u_i(g)^2 + sum g_i^2 u_i(q)^2  <=  |qbar|_4^2 W^2 + G4^2 |u(q)|_4^2
// (qbar = outward-rounded |q|; W = max_rows |u(g)|_4 and G4 = max_rows |g|_4 are recorded when the copy is built, w4max / g4max).
// The search is CERTIFIED per query (append_rerank_kernel / topk_rerank_kernel): with E = sqrt(V ln(1/p) / 2) + eps_det, rows are
// pruned only below  (best coarse) - m,  m = (1 + kF8GapFrac) E, and the result is accepted only if the best EXACT score L among the
// re-scored rows satisfies  L >= (best coarse) - m + E. If the true best row A had been pruned, then coarse_A < best coarse - m
// <= L - E <= exact_A - E: a single fixed row's error exceeded E, which has probability <= p = exp(-kF8LogP) (no union over the
// gallery). Queries that fail the certificate are recomputed by the exact fp32 scan. eps_det covers the deterministic parts: the
// tensor core's fp32 accumulation of the exact e4m3 products and the fp32 rounding of the exact re-score (kF8AccEps |qbar| |g|).
constexpr float kF8Scale = 256.f;
constexpr float kF8LogP = 27.631f;     // ln(1e12): per-query bound on the probability of a wrong top-1, for arbitrary rows / queries
constexpr float kF8GapFrac = 0.3f;     // room between the best coarse score and the best exact score before a query is recomputed
constexpr float kF8AccEps = 1.1e-3f;   // >= 2^-10 + gamma_512: accumulation in the MMA pipe and in dot512, relative to |qbar| |ghat|
constexpr float kF8Max = 448.f;        // largest e4m3 magnitude

// e4m3 neighbours of a scaled magnitude a in [0, 448]: lo = largest e4m3 value <= a, step = distance to the next one
// (2^(e-3) for a in [2^e, 2^(e+1)), e >= -6; 2^-9 in the subnormal range). Exact in fp32: power-of-two scalings and a floor.
__host__ __device__ __forceinline__ void f8_bracket(float a, float& lo, float& step) {
    union { float f; uint32_t u; } v;
    v.f = a;
    int e = static_cast<int>((v.u >> 23) & 0xFF) - 127;
    e = e < -6 ? -6 : e;
    v.u = static_cast<uint32_t>(e - 3 + 127) << 23;
    step = v.f;
#ifdef __CUDA_ARCH__
    lo = floorf(a / step) * step;
#else
    lo = static_cast<float>(static_cast<int>(a / step)) * step;
#endif
}
// Stochastic rounding of x * kF8Scale to e4m3 with the 24-bit uniform variate r24. Returns the rounded SCALED value (exactly
// representable); u = the step of the bracket (0 when the value is representable: no randomness), abar = outward-rounded magnitude.
__host__ __device__ __forceinline__ float f8_round_dither(float x, uint32_t r24, float& u, float& abar) {
    float a = x < 0.f ? -x : x;
    a *= kF8Scale;
    a = a > kF8Max ? kF8Max : a;  // saturation is reported by the callers (rows are refused, queries go to the exact scan)
    float lo, step;
    f8_bracket(a, lo, step);
    const float frac = (a - lo) / step;  // exact, in [0, 1)
    const bool inexact = frac > 0.f;
    const bool up = static_cast<float>(r24) < frac * 16777216.f;
    u = inexact ? step : 0.f;
    abar = inexact ? lo + step : lo;
    const float r = up ? lo + step : lo;
    return x < 0.f ? -r : r;
}
// "append" epilogue (top-1 searches on the fp8 scan copy): instead of a sorted register list every epilogue thread appends the
// rows that pass its running threshold to a private buffer in global memory; the wide fp8 margin makes passes frequent (a few per
// thousand rows), and an append is a predicated 8-byte store where a sorted insert is a divergent 8-deep compare-swap chain.
constexpr int kAppCap = 128;          // entries per (unit, column half, query); more than that hands the query to the exact scan
constexpr int kAppRescoreMax = 4096;  // rows re-scored exactly per query in append mode (unmatched queries on a 10 M-row shard: ~600)
template <int CG, bool F8 = false>
struct CoarseCfg {
    static constexpr int kKB = F8 ? 4 : 8;                         // k-blocks of 128 bytes per row
    static constexpr int kQSmem = kKB * kQTileBytes;               // 128 KiB (fp16) / 64 KiB (fp8)
    static constexpr int kRing = F8 ? 160 * 1024 : kRingBytes;     // the fp8 query block leaves room for a deeper ring
    static constexpr int kStageBytes = (kTileRows / CG) * 128;     // bytes this CTA loads per k-block
    static constexpr int kStages = kRing / kStageBytes;            // fp16: 3 / 6 (single CTA / pair); fp8: 5 / 10
    static constexpr int kSmemBytes = 1024 /*align slack*/ + kQSmem + kRing + 256 /*barriers*/;
};
// candidate list length per epilogue thread: KSEL = 1 (top-1 search) keeps 8, KSEL = 8 (k <= 8) keeps 16
template <int KSEL>
struct ListCfg {
    static constexpr int kKC = KSEL == 1 ? 8 : 16;
};

// ---- cross-GPU push of finished results (csrc/exchange_impl.cuh has the protocol): mailboxes of every rank, mapped into this one
struct __align__(16) XEntry {
    float score;
    int pad;
    long long idx;
};
struct XPeers {
    XEntry* entries[16];
    unsigned int* flags[16];
};
struct XPush {
    XPeers peers;
    int world, rank, nq_max, k_max;
    unsigned int* state;  // [0] pushes completed, [1] finished-block ticket of the running push
    int enabled;
};
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// number of the push this kernel performs = pushes completed + 1, read when the block starts (no block of the call can have
// finished the call yet). after_advance: an earlier kernel of the same search already advanced the counter (exact-scan fix-up).
__device__ __forceinline__ unsigned int xpush_epoch(const XPush& xp, bool after_advance) {
    const unsigned int e = *reinterpret_cast<volatile unsigned int*>(xp.state);
    return after_advance ? e : e + 1u;
}
// Every thread of the block calls. s_s / s_i: the query's k results (global rows) in shared memory, visible to the block.
__device__ __forceinline__ void xpush_query(const XPush& xp, unsigned int epoch, int q, int k, const float* s_s, const long long* s_i) {
    const int slot = epoch & 3;
    const size_t lane_base = (static_cast<size_t>(slot) * xp.world + xp.rank) * xp.nq_max + q;
    for (int t = threadIdx.x; t < xp.world * k; t += blockDim.x) {
        const int peer = t / k, j = t % k;
        XEntry e;
        e.score = s_s[j];
        e.pad = 0;
        e.idx = s_i[j];
        xp.peers.entries[peer][lane_base * xp.k_max + j] = e;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < xp.world) st_release_sys(xp.peers.flags[threadIdx.x] + lane_base, epoch);
}
// Once per block, at its end: the block that finishes last advances the push counter (after every block has read the old value).
__device__ __forceinline__ void xpush_finish(const XPush& xp, unsigned int epoch, unsigned int blocks) {
    if (threadIdx.x == 0 && atomicAdd(xp.state + 1, 1u) == blocks - 1) {
        xp.state[1] = 0u;
        __threadfence();
        xp.state[0] = epoch;
    }
}

__device__ __forceinline__ bool better(float sa, long long ia, float sb, long long ib) {
    return sa > sb || (sa == sb && ia < ib);
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {  // one FMNMX3
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// insert (v, id) into a descending list held in registers; ties keep the earlier (lower row) entry first
template <int KC>
__device__ __forceinline__ void topk_insert(float (&s)[KC], int (&ix)[KC], float v, int id) {
    s[KC - 1] = v;
    ix[KC - 1] = id;
#pragma unroll
    for (int t = KC - 1; t > 0; --t) {
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
// Epilogue warp e (0..7): TMEM lane quarter e & 3, accumulator column half e >> 2.
// cand_s / cand_i: [units * 2 lists][CG*128 queries][KC]  (coarse score, local row), descending, (-inf,-1) padded.
// A thread keeps a row iff its coarse score exceeds max(KC-th best, KSEL-th best - 2 eps |q| gmax): every row that can
// still reach the exact top-KSEL survives unless more than KC such rows exist (detected in topk_rerank_kernel).
// ----------------------------------------------------------------------------------------------------------
template <int CG, int KSEL, bool F8, bool APP = false>
__global__ void __launch_bounds__((coarse_threads<CG, F8, APP>()), 1)
cosine_topk_coarse(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapq, const float* __restrict__ q_margin,
                   int nq, long long n_rows, int num_tiles, float* __restrict__ cand_s, int* __restrict__ cand_i,
                   int* __restrict__ flag_list, int* __restrict__ gbest, uint2* __restrict__ app_buf, int* __restrict__ app_cnt) {
    static_assert(!APP || KSEL == 1, "the append epilogue serves top-1 searches");
    using Cfg = CoarseCfg<CG, F8>;
    constexpr int kKB = Cfg::kKB;
    if (blockIdx.x == 0 && threadIdx.x == 0) flag_list[0] = 0;  // list of queries the re-rank hands to the exact scan
    constexpr int KC = ListCfg<KSEL>::kKC;
    constexpr int kEW = coarse_epi_warps<CG, F8, APP>();  // epilogue warps
    constexpr int kSplit = kEW / 4;                       // column slices of an accumulator tile (one per epilogue warp group)
    constexpr int kColsPer = kTileRows / kSplit;
    static_assert(APP || kSplit == 2, "the sorted-list epilogue is written for two column halves");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* q_smem = smem;
    uint8_t* ring = smem + Cfg::kQSmem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + Cfg::kRing);  // [kStages]  (used in the leader CTA)
    uint64_t* empty_bar = full_bar + Cfg::kStages;                        // [kStages]
    uint64_t* tfull_bar = empty_bar + Cfg::kStages;                       // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;                                 // [2] accumulator drained (leader CTA)
    uint64_t* q_bar = tempty_bar + 2;                                     // queries staged in every CTA of the unit (leader CTA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int unit = blockIdx.x / CG;
    const int num_units = gridDim.x / CG;
    // Tile sequence of this unit: tiles unit, unit + num_units, ... . The e4m3 append scan DEFERS its first kDefer tiles: they are first
    // processed max-only (they seed the per-query threshold, own and - through gbest - every other unit's) and once more, with appends,
    // at the end of the unit's sequence. Before, a unit's first tile was appended against the threshold of its own 256 rows (best -
    // margin ~ 1 sigma: 16 % of its rows passed) and the second against little more: two tiles of 66 produced 65 % of all appends of a
    // 1.25 M-row shard (4600 per query for 1200 in-margin rows; ncu source counters). Cost: kDefer extra tiles of MMA work per unit.
    constexpr int kDefer = (APP && F8) ? FR_AB_DEFER : 0;
    const int n_my = unit < num_tiles ? (num_tiles - unit + num_units - 1) / num_units : 0;
    const int n_def = n_my < kDefer ? n_my : kDefer;
    const int n_it = n_my + n_def;
    auto tile_of = [&](int it) { return unit + (it < n_my ? it : it - n_my) * num_units; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap);
        tma_prefetch_desc(&tmapq);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], CG * kEW);  // one arrive per epilogue warp of every CTA of the unit
        }
        mbar_init(q_bar, 1);  // the leader's expect_tx arrival; both CTAs' query loads complete_tx on it
        fence_mbar_init();
    }
    if (warp == 2) {
        if (CG == 2) tmem_alloc_pair<512>(tmem_slot);
        else tmem_alloc<512>(tmem_slot);
    }

    // barriers and TMEM are ready in every CTA of the unit before anything is signalled across CTAs
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
            // this CTA's 128 queries: the operand image prep_queries_kernel wrote (fp16 / scaled e4m3, row-major 256 x 512), one
            // 128-row x 128-byte box per k-block, 128-byte swizzled by TMA like the gallery tiles
            if (cta_rank == 0) mbar_expect_tx(q_bar, CG * Cfg::kQSmem);
            for (int kb = 0; kb < kKB; ++kb) {
                if (CG == 1) tma_load_2d(q_smem + kb * kQTileBytes, &tmapq, q_bar, kb * (F8 ? 128 : 64), 0, kEvictLast);
                else tma_load_2d_pair(q_smem + kb * kQTileBytes, &tmapq, mapa_u32(smem_u32(q_bar), 0), kb * (F8 ? 128 : 64),
                                      static_cast<int>(cta_rank) * kQRows, kEvictLast);
            }
            for (int it = 0; it < n_it; ++it) {
                const int row0 = tile_of(it) * kTileRows;
                for (int kb = 0; kb < kKB; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* dst = ring + stage * Cfg::kStageBytes;
                    if (CG == 1) {
                        mbar_expect_tx(&full_bar[stage], 2 * kHalfTileBytes);
                        tma_load_2d(dst, &tmap, &full_bar[stage], kb * (F8 ? 128 : 64), row0, kEvictFirst);
                        tma_load_2d(dst + kHalfTileBytes, &tmap, &full_bar[stage], kb * (F8 ? 128 : 64), row0 + 128, kEvictFirst);
                    } else {
                        // the leader's barrier collects the bytes of both CTAs' halves
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kHalfTileBytes);
                        tma_load_2d_pair(dst, &tmap, leader_full0 + stage * 8, kb * (F8 ? 128 : 64), row0 + static_cast<int>(cta_rank) * 128,
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
            mbar_wait(q_bar, 0);  // query operand staged in both CTAs
            tc_fence_after();
            for (int it = 0; it < n_it; ++it) {
                const int buf = it & 1;
                mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * kTileRows;
                for (int kb = 0; kb < kKB; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(q_smem + kb * kQTileBytes);
                    const uint32_t b_addr = smem_u32(ring + stage * Cfg::kStageBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 4 MMAs (K = 16 fp16 / 32 fp8 = 32 bytes each) inside the 128-byte swizzle span
                        const uint64_t da = umma_desc_sw128(a_addr + k * 32);
                        const uint64_t db = umma_desc_sw128(b_addr + k * 32);
                        if (F8) {
                            if (CG == 2) umma_f8_ss_pair(d_tmem, da, db, idesc, (kb | k) != 0);
                            else umma_f8_ss(d_tmem, da, db, idesc, (kb | k) != 0);
                        } else {
                            if (CG == 2) umma_f16_ss_pair(d_tmem, da, db, idesc, (kb | k) != 0);
                            else umma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
                        }
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
        // ===================== epilogue: running candidate list per query, in registers =====================
        const int ew = warp & 3;          // TMEM lane quarter this warp may access (hardware: warp id % 4)
        const int half = (warp - 4) >> 2;  // accumulator column slice (half; quarter with sixteen epilogue warps)
        const int qrow = static_cast<int>(cta_rank) * kQRows + ew * 32 + lane;
        // margin = 2 eps |q| gmax in accumulator units (prep_queries_kernel); 0 for the padding rows beyond nq
        const float margin = qrow < nq ? __ldg(q_margin + qrow) : 0.f;
        if constexpr (APP) {
            // ---------- append epilogue (see kAppCap): thread state = running best, threshold, entry count
            constexpr float kRaw = F8 ? kF8Scale * kF8Scale : 1.f;
            constexpr float kInvRaw = 1.f / kRaw;
            constexpr int kChunksA = kColsPer / kLdCols;
            const bool live = qrow < nq;
            const size_t list = (static_cast<size_t>(unit) * kSplit + half) * (CG * kQRows) + qrow;
            uint2* mybuf = app_buf + list * kAppCap;
            float best = -INFINITY, thr = live ? -INFINITY : INFINITY, published = 0.f;
            int cnt = 0;
            const uint32_t tempty_leader = (CG == 2) ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;
            volatile int* gb_ptr = reinterpret_cast<volatile int*>(gbest + (live ? qrow : 0));

            // max of a chunk; tri[g] = max of columns 3g .. 3g+2 (g < 5), column 15 stands alone
            auto chunk_max = [&](const uint32_t (&raw)[kLdCols], int col0, int valid, float (&v)[kLdCols], float (&tri)[5]) -> float {
#pragma unroll
                for (int j = 0; j < kLdCols; ++j) v[j] = __uint_as_float(raw[j]);
                if (valid < kTileRows) {
#pragma unroll
                    for (int j = 0; j < kLdCols; ++j)
                        if (col0 + j >= valid) v[j] = -INFINITY;
                }
#pragma unroll
                for (int g = 0; g < 5; ++g) tri[g] = fmax3(v[3 * g], v[3 * g + 1], v[3 * g + 2]);
                return fmax3(fmax3(tri[0], tri[1], tri[2]), fmax3(tri[3], tri[4], v[15]), -INFINITY);
            };
            // Under the e4m3 copy's wide margin about a quarter of all (warp, chunk) pairs hold a passing value in SOME lane (ncu source
            // page, 1.25 M-row shard, queries without a match): the straight 16-way predicated append was 240 instructions and 45 % of the
            // epilogue's time. The slow path now descends by warp votes (uniform branches): column triples first, then single columns, so
            // the usual case (one lane, one column) runs one append body.
            auto append_if = [&](float val, int row) {
                if (val > thr) {
                    if (cnt < kAppCap) mybuf[cnt] = make_uint2(__float_as_uint(val * kInvRaw), static_cast<uint32_t>(row));
                    ++cnt;
                }
            };
            auto consume_app = [&](const uint32_t (&raw)[kLdCols], int col0, int valid, int row_base) {
                float v[kLdCols], tri[5];
                const float m = chunk_max(raw, col0, valid, v, tri);
                if (__any_sync(0xffffffffu, m > thr)) {
#pragma unroll
                    for (int g = 0; g < 5; ++g) {
                        if (__any_sync(0xffffffffu, tri[g] > thr)) {
#pragma unroll
                            for (int j = 3 * g; j < 3 * g + 3; ++j)
                                if (__any_sync(0xffffffffu, v[j] > thr)) append_if(v[j], row_base + col0 + j);
                        }
                    }
                    if (__any_sync(0xffffffffu, v[15] > thr)) append_if(v[15], row_base + col0 + 15);
                    best = fmaxf(best, m);
                    thr = fmaxf(thr, best - margin);
                }
            };

            int gb_bits = 0;  // the shared best of this query as fetched during the previous tile (bits of a non-negative float)
            for (int it = 0; it < n_it; ++it) {
                const int t = tile_of(it);
                const int buf = it & 1;
                mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
                tc_fence_after();
                // the shared best: the value fetched a whole tile ago is applied now and the next fetch is issued, so its L2 round trip
                // hides behind this tile AND the wait for the next accumulator (with sixteen epilogue warps a warp's share of a tile is
                // shorter than the round trip: consumed at the end of the same tile it was 14 % of the epilogue's time)
                if (FR_AB_GBLAG && live) {
                    const float gb = __int_as_float(gb_bits);
                    if (gb > 0.f) thr = fmaxf(thr, gb * kRaw - margin);
                }
                gb_bits = *gb_ptr;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * kTileRows + half * kColsPer;
                const long long row0 = static_cast<long long>(t) * kTileRows;
                const int valid = (n_rows - row0 >= kTileRows) ? kTileRows : static_cast<int>(n_rows - row0);
                const int row_base = static_cast<int>(row0);
                const int cbase = half * kColsPer;
                uint32_t ra[kLdCols], rb[kLdCols];
                const bool max_only = it < n_def;  // deferred tile, first visit
                if (max_only || (kDefer == 0 && it == 0)) {
                    // a max-only pass seeds the threshold (and the shared best) before anything is appended; without it the first chunks
                    // would append every row they see. kDefer == 0: the first tile is read twice (this pass, then the append pass below).
                    float tm = -INFINITY;
                    float v[kLdCols], tri[5];
                    tmem_ld_32x32b_x16(taddr, ra);
#pragma unroll 1
                    for (int c = 0; c < kChunksA; c += 2) {
                        tmem_ld_wait_x16(ra);
                        tmem_ld_32x32b_x16(taddr + (c + 1) * kLdCols, rb);
                        tm = fmaxf(tm, chunk_max(ra, cbase + c * kLdCols, valid, v, tri));
                        tmem_ld_wait_x16(rb);
                        if (c + 2 < kChunksA) tmem_ld_32x32b_x16(taddr + (c + 2) * kLdCols, ra);
                        tm = fmaxf(tm, chunk_max(rb, cbase + (c + 1) * kLdCols, valid, v, tri));
                    }
                    if (live) {
                        // strictly below the tile's best, so that the append pass keeps it
                        thr = fmaxf(thr, fminf(tm - margin, tm - fabsf(tm) * 1e-6f - 1e-30f));
                        if (tm > 0.f) atomicMax(gbest + qrow, __float_as_int(tm * kInvRaw));
                    }
                }
                if (!max_only) {
                    tmem_ld_32x32b_x16(taddr, ra);
#pragma unroll 1
                    for (int c = 0; c < kChunksA; c += 2) {
                        tmem_ld_wait_x16(ra);
                        tmem_ld_32x32b_x16(taddr + (c + 1) * kLdCols, rb);
                        consume_app(ra, cbase + c * kLdCols, valid, row_base);
                        tmem_ld_wait_x16(rb);
                        if (c + 2 < kChunksA) tmem_ld_32x32b_x16(taddr + (c + 2) * kLdCols, ra);
                        consume_app(rb, cbase + (c + 1) * kLdCols, valid, row_base);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_cluster(tempty_leader + buf * 8);
                    else mbar_arrive(&tempty_bar[buf]);
                }
                if (live) {
                    if (best > published && best > 0.f) {
                        published = best;
                        atomicMax(gbest + qrow, __float_as_int(best * kInvRaw));  // non-negative floats order like their bits
                    }
                    // last deferred tile: the appends start with the next one, so read the shared best NOW (one exposed L2 round trip):
                    // every unit published its first tile a whole tile ago
                    if (kDefer > 0 && it == n_def - 1) {
                        const float gb = __int_as_float(*gb_ptr);
                        if (gb > 0.f) thr = fmaxf(thr, gb * kRaw - margin);
                    }
                    if (!FR_AB_GBLAG) {  // A/B: the fetch of this tile applied at its own end
                        const float gb = __int_as_float(gb_bits);
                        if (gb > 0.f) thr = fmaxf(thr, gb * kRaw - margin);
                    }
                }
            }
            // [query][list] so that the re-rank reads one query's counts contiguously
            app_cnt[static_cast<size_t>(qrow) * (kSplit * num_units) + unit * kSplit + half] = live ? cnt : 0;
        } else {
        float best_s[KC];
        int best_i[KC];
#pragma unroll
        for (int j = 0; j < KC; ++j) {
            best_s[j] = -INFINITY;
            best_i[j] = -1;
        }
        float thr = -INFINITY;
        // top-1 searches share the best coarse score seen so far for each query between ALL epilogue threads of the GPU (gbest, cosine
        // units, as int bits of a positive float; reset by the re-rank kernel): any row's score bounds the final best from below, so
        // rows more than the margin below it can never be needed. It turns the per-thread list threshold into a global one.
        constexpr float kRawScale = F8 ? kF8Scale * kF8Scale : 1.f;
        float published = 0.f;
        const uint32_t tempty_leader0 = (CG == 2) ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;
        constexpr int kChunks = (kTileRows / 2) / kLdCols;  // 8 loads of 16 columns per tile half

        auto consume = [&](const uint32_t (&raw)[kLdCols], int col0, int valid, int row_base) {
            float v[kLdCols];
#pragma unroll
            for (int j = 0; j < kLdCols; ++j) v[j] = __uint_as_float(raw[j]);
            if (valid < kTileRows) {  // last, partial tile: rows beyond the gallery are TMA zero fill
#pragma unroll
                for (int j = 0; j < kLdCols; ++j)
                    if (col0 + j >= valid) v[j] = -INFINITY;
            }
            float m[kLdCols / 2];
#pragma unroll
            for (int j = 0; j < kLdCols / 2; ++j) m[j] = fmaxf(v[2 * j], v[2 * j + 1]);
#pragma unroll
            for (int w = kLdCols / 4; w > 0; w >>= 1)
#pragma unroll
                for (int j = 0; j < w; ++j) m[j] = fmaxf(m[j], m[j + w]);
            if (m[0] > thr) {
#pragma unroll
                for (int j = 0; j < kLdCols; ++j) {
                    if (v[j] > thr) {
                        topk_insert<KC>(best_s, best_i, v[j], row_base + col0 + j);
                        thr = fmaxf(best_s[KC - 1], best_s[KSEL - 1] - margin);
                    }
                }
            }
        };

        for (int it = 0; it < n_it; ++it) {
            const int t = tile_of(it);
            const int buf = it & 1;
            mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * kTileRows + half * (kTileRows / 2);
            const long long row0 = static_cast<long long>(t) * kTileRows;
            const int valid = (n_rows - row0 >= kTileRows) ? kTileRows : static_cast<int>(n_rows - row0);
            const int row_base = static_cast<int>(row0);
            const int cbase = half * (kTileRows / 2);
            if (KSEL == 1 && qrow < nq) {
                const float gb = __int_as_float(*reinterpret_cast<volatile int*>(gbest + qrow));
                if (gb > 0.f) thr = fmaxf(thr, gb * kRawScale - margin);
            }
            // software pipeline: the load of chunk c+1 is in flight while chunk c is consumed
            uint32_t ra[kLdCols], rb[kLdCols];
            tmem_ld_32x32b_x16(taddr, ra);
#pragma unroll 1
            for (int c = 0; c < kChunks; c += 2) {
                tmem_ld_wait_x16(ra);
                tmem_ld_32x32b_x16(taddr + (c + 1) * kLdCols, rb);
                consume(ra, cbase + c * kLdCols, valid, row_base);
                tmem_ld_wait_x16(rb);
                if (c + 2 < kChunks) tmem_ld_32x32b_x16(taddr + (c + 2) * kLdCols, ra);
                consume(rb, cbase + (c + 1) * kLdCols, valid, row_base);
            }
            tc_fence_before();
            __syncwarp();
            if (KSEL == 1 && qrow < nq && best_s[0] > published) {
                published = best_s[0];
                atomicMax(gbest + qrow, __float_as_int(best_s[0] * (1.f / kRawScale)));  // non-negative floats order like their bits
            }
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(tempty_leader0 + buf * 8);
                else mbar_arrive(&tempty_bar[buf]);
            }
        }
        const size_t o = ((static_cast<size_t>(unit) * 2 + half) * (CG * kQRows) + qrow) * KC;
#pragma unroll
        for (int j = 0; j < KC; ++j) {
            cand_s[o + j] = F8 ? best_s[j] * (1.f / (kF8Scale * kF8Scale)) : best_s[j];
            cand_i[o + j] = best_i[j];
        }
        }  // !APP
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
// so that sims, re-rank and exact scan agree bit for bit).  a: 16 registers per lane, element (lane + 32 i) * 4 + e.
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
constexpr int kHeadMax = 2 * 148 * kTopkMax;  // first k entries of every candidate list
constexpr int kRescoreMax = 64;               // rows re-scored exactly per query before the query is declared "overflowed"

// One block per query. lists x kc coarse candidates -> exact top-k, or flag[q] = 1 when the candidate set may be incomplete
// (more than kRescoreMax rows inside the margin, or a list that was full of rows inside the margin).
// out_s: nq x k, out_i: nq x k (row_offset + local row), padded with (-inf, -1).
__global__ void __launch_bounds__(kSelThreads) topk_rerank_kernel(const float* __restrict__ cand_s, const int* __restrict__ cand_i,
                                                                  int lists, int q_stride, int kc, const float* __restrict__ q,
                                                                  const float* __restrict__ rows, const float* __restrict__ q_margin,
                                                                  const float* __restrict__ q_gap, float inv_raw, int k, long long row_offset,
                                                                  float* __restrict__ out_s, long long* __restrict__ out_i,
                                                                  int* __restrict__ flag_list, int* __restrict__ gbest, const XPush push) {
    __shared__ float cs[kHeadMax];
    __shared__ long long x_i[kTopkMax];
    const unsigned int push_epoch = push.enabled ? xpush_epoch(push, false) : 0u;
    __shared__ long long ci[kHeadMax];
    __shared__ float sel_s[kTopkMax];
    __shared__ long long sel_i[kTopkMax];
    __shared__ float red_s[32];
    __shared__ long long red_i[32];
    __shared__ int red_p[32];
    __shared__ float rs[kRescoreMax];
    __shared__ long long ri[kRescoreMax];
    __shared__ int n_resc, overflow;
    const int qi = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        n_resc = 0;
        overflow = 0;
    }
    // phase 1: k-th best coarse score over all lists (lists are sorted, so only their first k entries matter)
    const int head = lists * k;
    for (int p = threadIdx.x; p < head; p += blockDim.x) {
        const int l = p / k, j = p % k;
        const size_t o = (static_cast<size_t>(l) * q_stride + qi) * kc + j;
        cs[p] = cand_s[o];
        ci[p] = cand_i[o];
    }
    float4 qa[4];
    load512(q + static_cast<size_t>(qi) * kDim, lane, qa);
    __syncthreads();
    block_select(cs, ci, head, k, sel_s, sel_i, red_s, red_i, red_p);
    const float ck = sel_s[k - 1];  // -inf when fewer than k rows exist
    const float thr = ck - __ldg(q_margin + qi) * inv_raw;  // the scan's margin (prep_queries_kernel), in cosine units
    // phase 2: every candidate with coarse >= thr is re-scored
    const int total = lists * kc;
    for (int p = threadIdx.x; p < total; p += blockDim.x) {
        const int l = p / kc, j = p % kc;
        const size_t o = (static_cast<size_t>(l) * q_stride + qi) * kc + j;
        const int id = cand_i[o];
        if (id < 0) continue;
        const float s = cand_s[o];
        if (s >= thr) {
            if (j == kc - 1) overflow = 1;  // this list was full of in-margin rows: it may have dropped one
            const int slot = atomicAdd(&n_resc, 1);
            if (slot < kRescoreMax) ri[slot] = id;
        }
    }
    __syncthreads();
    const int nr = min(n_resc, kRescoreMax);
    if (n_resc > kRescoreMax) overflow = 1;
    // phase 3: exact scores
    for (int c = warp; c < nr; c += (blockDim.x >> 5)) {
        float4 b[4];
        load512(rows + static_cast<size_t>(ri[c]) * kDim, lane, b);
        const float s = dot512(qa, b);
        if (lane == 0) rs[c] = s;
    }
    __syncthreads();
    // phase 4: final order by (exact score desc, row asc)
    block_select(rs, ri, nr, k, sel_s, sel_i, red_s, red_i, red_p);
    if (threadIdx.x < k) {
        const long long id = sel_i[threadIdx.x];
        out_s[static_cast<size_t>(qi) * k + threadIdx.x] = sel_s[threadIdx.x];
        out_i[static_cast<size_t>(qi) * k + threadIdx.x] = id >= 0 ? id + row_offset : -1;
    }
    // certificate of the e4m3 scan (see kF8LogP): the k-th best EXACT score must not fall more than the gap below the k-th best
    // coarse score, else a pruned row could belong to the top-k with probability > p. (+inf gap on the fp16 copy: always passes.)
    // (a gap of -inf marks a query whose operand image is outside the scan copy's range: always recomputed, even when its coarse
    // scores were NaN and nothing was kept)
    if (threadIdx.x == 0 && (__ldg(q_gap + qi) == -INFINITY || sel_s[k - 1] < ck - __ldg(q_gap + qi))) overflow = 1;
    if (threadIdx.x == 0 && overflow) flag_list[1 + atomicAdd(&flag_list[0], 1)] = qi;
    if (threadIdx.x == 0) gbest[qi] = 0;  // ready for the next search (0 = nothing published)
    if (push.enabled) {  // deliver this query to every peer now, unless the exact scan is going to recompute (and deliver) it
        if (threadIdx.x < k) x_i[threadIdx.x] = sel_i[threadIdx.x] >= 0 ? sel_i[threadIdx.x] + row_offset : -1;
        __syncthreads();
        if (!overflow) xpush_query(push, push_epoch, qi, k, sel_s, x_i);
        xpush_finish(push, push_epoch, gridDim.x);
    }
}

// Re-rank for the append epilogue (top-1). One block per query.
//  1. thread t walks lists t, t + 256 (count <= kAppCap entries each, contiguous): entries within the scan's margin m of the best coarse
//     score (the scan's shared per-query best, or a pass over the entries when nothing positive was published) are collected.
//  2. e4m3 copy only (finite gap): the eight per-warp coarse leaders are re-scored exactly, L0 = the best of them, and only rows with
//     coarse >= L0 - E stay. Sound under the SAME event as the certificate: the true best A has exact_A >= L0, so dropping it here
//     means exact_A - coarse_A > E, the one-row event of probability <= p (kF8LogP) the certificate already charges. It cuts the
//     candidates of a query without a match ~4x (m = 1.3 E below the best COARSE score -> E below an EXACT one).
//  3. more than 32 rows left: pre-filter on the fp16 copy (deterministic error, see below); 4. exact fp32 scores, best by (score
//     desc, row asc), certificate, push to the peers.
// A list that overflowed (count > kAppCap) or more than kAppRescoreMax in-margin rows flag the query for the exact scan.
__global__ void __launch_bounds__(kSelThreads, 2) append_rerank_kernel(const uint2* __restrict__ app_buf, const int* __restrict__ app_cnt,
                                                                    int lists, int q_stride, const float* __restrict__ q,
                                                                    const float* __restrict__ rows, const float* __restrict__ q_margin,
                                                                    const float* __restrict__ q_gap, float inv_raw, long long row_offset,
                                                                    float* __restrict__ out_s, long long* __restrict__ out_i,
                                                                    int* __restrict__ flag_list, int* __restrict__ gbest, const XPush push,
                                                                    const __half* __restrict__ rows_f16, const float* __restrict__ gmax) {
    constexpr int kW = kSelThreads / 32;
    constexpr int kPer = kAppRescoreMax / kSelThreads;
    constexpr int kListsPer = (2 * 148 + kSelThreads - 1) / kSelThreads;  // lists per thread (lists <= kMaxLists = 296)
    __shared__ float x_s[1];
    __shared__ long long x_i[1];
    const unsigned int push_epoch = push.enabled ? xpush_epoch(push, false) : 0u;
    __shared__ float rs[kAppRescoreMax];
    __shared__ int ri[kAppRescoreMax];
    __shared__ float red_s[kW];
    __shared__ int red_i[kW];
    __shared__ int n_resc, overflow;
    const int qi = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        n_resc = 0;
        overflow = 0;
    }
    // this thread's lists: entry pointers and (clamped) counts
    const uint2* lp[kListsPer];
    int lc[kListsPer];
#pragma unroll
    for (int j = 0; j < kListsPer; ++j) {
        const int l = threadIdx.x + j * kSelThreads;
        lc[j] = l < lists ? app_cnt[static_cast<size_t>(qi) * lists + l] : 0;  // [query][list]: one contiguous read per query
        lp[j] = app_buf + (static_cast<size_t>(l < lists ? l : 0) * q_stride + qi) * kAppCap;
    }
    float4 qa[4];
    load512(q + static_cast<size_t>(qi) * kDim, lane, qa);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kListsPer; ++j)
        if (lc[j] > kAppCap) {
            overflow = 1;
            lc[j] = kAppCap;
        }
    auto block_max = [&](float mx) {  // all threads; leaves the result in every thread
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        __syncthreads();  // red_s may still be read from a previous use
        if (lane == 0) red_s[warp] = mx;
        __syncthreads();
        mx = red_s[0];
#pragma unroll
        for (int w = 1; w < kW; ++w) mx = fmaxf(mx, red_s[w]);
        return mx;
    };
    float ck = __int_as_float(gbest[qi]);  // cosine units; 0 = nothing positive was published
    if (!(ck > 0.f)) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kListsPer; ++j)
            for (int e = 0; e < lc[j]; ++e) mx = fmaxf(mx, __uint_as_float(__ldg(lp[j] + e).x));
        ck = block_max(mx);
    }
    const float m_cos = __ldg(q_margin + qi) * inv_raw;  // the scan's margin (prep_queries_kernel), in cosine units
    const float thr = ck - m_cos;
#pragma unroll
    for (int j = 0; j < kListsPer; ++j) {
        for (int e0 = 0; e0 < lc[j]; e0 += 4) {  // four independent loads in flight (a list is contiguous: mostly one or two lines)
            uint2 en[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) en[t] = __ldg(lp[j] + (e0 + t < lc[j] ? e0 + t : e0));
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (e0 + t < lc[j] && __uint_as_float(en[t].x) >= thr) {
                    const int slot = atomicAdd(&n_resc, 1);
                    if (slot < kAppRescoreMax) {
                        rs[slot] = __uint_as_float(en[t].x);
                        ri[slot] = static_cast<int>(en[t].y);
                    }
                }
            }
        }
    }
    __syncthreads();
    int nr = min(n_resc, kAppRescoreMax);
    if (n_resc > kAppRescoreMax && threadIdx.x == 0) overflow = 1;
    const bool had_rows = nr > 0;
    // in-place compaction of ri[0, nr): keeps the rows whose rs[] value is >= keep_thr. Every thread reads its candidates before the
    // barrier, survivors are re-appended after it (rs[] is stale afterwards: the next phase rewrites it).
    auto compact = [&](float keep_thr) {
        int keep_id[kPer];
        int n_keep = 0;
#pragma unroll
        for (int t = 0; t < kPer; ++t) {
            const int c = threadIdx.x + t * kSelThreads;
            keep_id[t] = (c < nr && rs[c] >= keep_thr) ? ri[c] : -1;
            n_keep += keep_id[t] >= 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) n_resc = 0;
        __syncthreads();
        if (n_keep) {
            int slot = atomicAdd(&n_resc, n_keep);
#pragma unroll
            for (int t = 0; t < kPer; ++t)
                if (keep_id[t] >= 0) ri[slot++] = keep_id[t];
        }
        __syncthreads();
        nr = n_resc;
    };
    const float gap = __ldg(q_gap + qi);
    if (gap > 0.f && gap < INFINITY && nr > kW) {
        // e4m3 copy: the per-warp coarse leaders (disjoint subsets; the overall coarse leader is one of them), exactly re-scored
        float bs = -INFINITY;
        int bc = -1;
        for (int c = threadIdx.x; c < nr; c += kSelThreads)
            if (rs[c] > bs) {
                bs = rs[c];
                bc = c;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (oc >= 0 && (bc < 0 || os > bs)) {
                bs = os;
                bc = oc;
            }
        }
        bc = __shfl_sync(0xffffffffu, bc, 0);  // equal scores may leave different lanes with different slots: lane 0 decides
        float l0 = -INFINITY;
        if (bc >= 0) {  // warp-uniform
            float4 b[4];
            load512(rows + static_cast<size_t>(ri[bc]) * kDim, lane, b);
            l0 = dot512(qa, b);
        }
        l0 = block_max(l0);
        const float E = m_cos - gap;  // m = (1 + kF8GapFrac) E, gap = kF8GapFrac E
        compact(fmaxf(thr, l0 - E));
    }
    // Pre-filter on the fp16 copy (e4m3 scan with many in-margin rows): an fp16 row costs half the bytes of an fp32 row and its score
    // carries a DETERMINISTIC error |s16 - exact| <= eps16 = kCoarseEps |q| gmax, so only rows with s16 >= max s16 - 2 eps16 can hold the
    // exact maximum: the true best A satisfies s16_A >= exact_A - eps16 >= exact_B - eps16 >= s16_B - 2 eps16 for the fp16 leader B.
    // No probability is involved; the certificate below still judges the exact score.
    if (rows_f16 != nullptr && nr > 32) {
        float q16[16];  // the query in the fp16 row's lane layout: elements 8 (lane + 32 j) .. + 7, j = 0, 1
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float4* qp = reinterpret_cast<const float4*>(q + static_cast<size_t>(qi) * kDim) + 2 * (lane + 32 * j);
            const float4 a = __ldg(qp), b = __ldg(qp + 1);
            q16[8 * j + 0] = a.x, q16[8 * j + 1] = a.y, q16[8 * j + 2] = a.z, q16[8 * j + 3] = a.w;
            q16[8 * j + 4] = b.x, q16[8 * j + 5] = b.y, q16[8 * j + 6] = b.z, q16[8 * j + 7] = b.w;
        }
        const float qn = sqrtf(dot512(qa, qa));
        const float eps16 = kCoarseEps * qn * __ldg(gmax);
        auto dot16 = [&](const uint4& v0, const uint4& v1) {
            float acc = 0.f;
            const __half2* h0 = reinterpret_cast<const __half2*>(&v0);
            const __half2* h1 = reinterpret_cast<const __half2*>(&v1);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 f0 = __half22float2(h0[t]), f1 = __half22float2(h1[t]);
                acc = fmaf(q16[2 * t], f0.x, acc);
                acc = fmaf(q16[2 * t + 1], f0.y, acc);
                acc = fmaf(q16[8 + 2 * t], f1.x, acc);
                acc = fmaf(q16[8 + 2 * t + 1], f1.y, acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            return acc;
        };
        constexpr int kFly = 6;  // rows in flight per warp (the gather is latency-bound: 1 KiB per row, 8 warps per block, 2 blocks per SM)
        for (int c = warp; c < nr; c += kFly * kW) {
            uint4 v[kFly][2];
#pragma unroll
            for (int t = 0; t < kFly; ++t) {
                const int ct = c + t * kW;
                const uint4* r = reinterpret_cast<const uint4*>(rows_f16 + static_cast<size_t>(ri[ct < nr ? ct : c]) * kDim);
                v[t][0] = __ldg(r + lane);
                v[t][1] = __ldg(r + lane + 32);
            }
#pragma unroll
            for (int t = 0; t < kFly; ++t) {
                const float s = dot16(v[t][0], v[t][1]);
                if (lane == 0 && c + t * kW < nr) rs[c + t * kW] = s;
            }
        }
        __syncthreads();
        float mx = -INFINITY;
        for (int c = threadIdx.x; c < nr; c += kSelThreads) mx = fmaxf(mx, rs[c]);
        mx = block_max(mx);
        compact(mx - 2.f * eps16);
    }
    for (int c = warp; c < nr; c += 2 * kW) {  // two rows in flight per warp
        const int c2 = c + kW;
        float4 b0[4], b1[4];
        load512(rows + static_cast<size_t>(ri[c]) * kDim, lane, b0);
        load512(rows + static_cast<size_t>(ri[c2 < nr ? c2 : c]) * kDim, lane, b1);
        const float s0 = dot512(qa, b0), s1 = dot512(qa, b1);
        if (lane == 0) {
            rs[c] = s0;
            if (c2 < nr) rs[c2] = s1;
        }
    }
    __syncthreads();
    // best by (exact score desc, row asc)
    float bs = -INFINITY;
    int bi = -1;
    for (int c = threadIdx.x; c < nr; c += kSelThreads) {
        const float sc = rs[c];
        const int id = ri[c];
        if (bi < 0 || sc > bs || (sc == bs && id < bi)) {
            bs = sc;
            bi = id;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) {
            bs = os;
            bi = oi;
        }
    }
    if (lane == 0) {
        red_s[warp] = bs;
        red_i[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kW; ++w) {
            const float os = red_s[w];
            const int oi = red_i[w];
            if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) {
                bs = os;
                bi = oi;
            }
        }
        out_s[qi] = bi >= 0 ? bs : -INFINITY;
        out_i[qi] = bi >= 0 ? bi + row_offset : -1;
        // certificate of the e4m3 scan (see kF8LogP): accept only if the best exact score is within the gap of the best coarse score;
        // then a pruned true best would need a rounding error beyond E. (+inf gap on the fp16 copy: always passes.)
        // (a gap of -inf marks a query outside the scan copy's range: always recomputed, even if its coarse scores were NaN)
        // (had_rows && bi < 0: the leader filter emptied the set, which needs the same rare event: recompute)
        if (gap == -INFINITY || (bi >= 0 && bs < ck - gap) || (bi < 0 && had_rows)) overflow = 1;
        if (overflow) flag_list[1 + atomicAdd(&flag_list[0], 1)] = qi;
        gbest[qi] = 0;  // ready for the next search
        x_s[0] = bi >= 0 ? bs : -INFINITY;
        x_i[0] = bi >= 0 ? bi + row_offset : -1;
    }
    if (push.enabled) {  // deliver this query to every peer now, unless the exact scan is going to recompute (and deliver) it
        __syncthreads();
        if (!overflow) xpush_query(push, push_epoch, qi, 1, x_s, x_i);
        xpush_finish(push, push_epoch, gridDim.x);
    }
}

// exact fp32 scan. flag_list == nullptr: all nq queries; else the flag_list[0] queries listed in flag_list[1..].
// grid (slices, qsplit), 256 threads: block (s, y) handles queries y, y + qsplit, ...; warp w of slice s scores rows
// s*8+w, s*8+w + 8*slices, ... and keeps its top-k; the block merges its 8 warps.  part_s / part_i: [nq][slices][kTopkMax]
// Fix-up mode (flag_list != nullptr, grid.y == 1): nothing flagged -> every block returns at once; otherwise the block that finishes
// last (ticket) also merges the slices of the flagged queries into out_s / out_i, so the fix-up is ONE launch per search.
constexpr int kScanThreads = 256;
constexpr int kScanSlicesMax = 148 * 2;
__global__ void __launch_bounds__(kScanThreads) exact_scan_kernel(const float* __restrict__ rows, long long n, const float* __restrict__ q,
                                                                  int nq, const int* __restrict__ flag_list, float* part_s, long long* part_i,
                                                                  int k, long long row_offset, float* __restrict__ out_s,
                                                                  long long* __restrict__ out_i, unsigned int* __restrict__ ticket,
                                                                  int* __restrict__ flagged_acc, const XPush push) {
    const int total = flag_list ? flag_list[0] : nq;
    if (total == 0) return;
    if (flagged_acc && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(flagged_acc, total);
    for (int f = blockIdx.y; f < total; f += gridDim.y) {
    const int qi = flag_list ? flag_list[1 + f] : f;
    __shared__ float cs[8 * kTopkMax];
    __shared__ long long ci[8 * kTopkMax];
    __shared__ float sel_s[kTopkMax];
    __shared__ long long sel_i[kTopkMax];
    __shared__ float red_s[32];
    __shared__ long long red_i[32];
    __shared__ int red_p[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 qa[4];
    load512(q + static_cast<size_t>(qi) * kDim, lane, qa);
    float bs[kTopkMax];
    int bi[kTopkMax];
#pragma unroll
    for (int j = 0; j < kTopkMax; ++j) {
        bs[j] = -INFINITY;
        bi[j] = -1;
    }
    const long long stride = static_cast<long long>(gridDim.x) * 8;
    for (long long r = static_cast<long long>(blockIdx.x) * 8 + warp; r < n; r += stride) {
        float4 b[4];
        load512(rows + static_cast<size_t>(r) * kDim, lane, b);
        const float s = dot512(qa, b);  // identical in every lane
        if ((bi[kTopkMax - 1] < 0 || s > bs[kTopkMax - 1]) && s == s) topk_insert<kTopkMax>(bs, bi, s, static_cast<int>(r));
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < kTopkMax; ++j) {
            cs[warp * kTopkMax + j] = bs[j];
            ci[warp * kTopkMax + j] = bi[j];
        }
    }
    __syncthreads();
    block_select(cs, ci, 8 * kTopkMax, kTopkMax, sel_s, sel_i, red_s, red_i, red_p);
    if (threadIdx.x < kTopkMax) {
        const size_t o = (static_cast<size_t>(qi) * gridDim.x + blockIdx.x) * kTopkMax + threadIdx.x;
        part_s[o] = sel_s[threadIdx.x];
        part_i[o] = sel_i[threadIdx.x];
    }
    __syncthreads();
    }
    if (!flag_list) return;
    // fix-up mode: last block merges
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        if (is_last) *ticket = 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    __shared__ float ms[kScanSlicesMax * kTopkMax];
    __shared__ long long mi[kScanSlicesMax * kTopkMax];
    __shared__ float msel_s[kTopkMax];
    __shared__ long long msel_i[kTopkMax];
    __shared__ float mred_s[32];
    __shared__ long long mred_i[32];
    __shared__ int mred_p[32];
    const int slices = gridDim.x, count = slices * kTopkMax;
    for (int f = 0; f < total; ++f) {
        const int qi = flag_list[1 + f];
        for (int p = threadIdx.x; p < count; p += blockDim.x) {
            ms[p] = __ldcg(part_s + static_cast<size_t>(qi) * count + p);
            mi[p] = __ldcg(part_i + static_cast<size_t>(qi) * count + p);
        }
        __syncthreads();
        block_select(ms, mi, count, k, msel_s, msel_i, mred_s, mred_i, mred_p);
        if (threadIdx.x < k) {
            const long long id = msel_i[threadIdx.x];
            out_s[static_cast<size_t>(qi) * k + threadIdx.x] = msel_s[threadIdx.x];
            out_i[static_cast<size_t>(qi) * k + threadIdx.x] = id >= 0 ? id + row_offset : -1;
            msel_i[threadIdx.x] = id >= 0 ? id + row_offset : -1;
        }
        __syncthreads();
        // the re-rank kernel skipped this query's push (and already advanced the push counter): deliver the recomputed result
        if (push.enabled) xpush_query(push, xpush_epoch(push, true), qi, k, msel_s, msel_i);
        __syncthreads();
    }
}
// second half of the exact scan (FR_PATH_EXACT: all queries): merge the slices of each query and write the final nq x k result
__global__ void __launch_bounds__(kSelThreads) exact_merge_kernel(const float* __restrict__ part_s, const long long* __restrict__ part_i,
                                                                  int slices, int nq, const int* __restrict__ flag_list, int k,
                                                                  long long row_offset, float* __restrict__ out_s,
                                                                  long long* __restrict__ out_i) {
    const int total = flag_list ? flag_list[0] : nq;
    for (int f = blockIdx.x; f < total; f += gridDim.x) {
    const int qi = flag_list ? flag_list[1 + f] : f;
    __shared__ float cs[kScanSlicesMax * kTopkMax];
    __shared__ long long ci[kScanSlicesMax * kTopkMax];
    __shared__ float sel_s[kTopkMax];
    __shared__ long long sel_i[kTopkMax];
    __shared__ float red_s[32];
    __shared__ long long red_i[32];
    __shared__ int red_p[32];
    const int count = slices * kTopkMax;
    for (int p = threadIdx.x; p < count; p += blockDim.x) {
        cs[p] = part_s[static_cast<size_t>(qi) * count + p];
        ci[p] = part_i[static_cast<size_t>(qi) * count + p];
    }
    __syncthreads();
    block_select(cs, ci, count, k, sel_s, sel_i, red_s, red_i, red_p);
    if (threadIdx.x < k) {
        const long long id = sel_i[threadIdx.x];
        out_s[static_cast<size_t>(qi) * k + threadIdx.x] = sel_s[threadIdx.x];
        out_i[static_cast<size_t>(qi) * k + threadIdx.x] = id >= 0 ? id + row_offset : -1;
    }
    __syncthreads();
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

// merge of n_parts per-shard results (each nq x k, global indices) -> nq x k, order (score desc, idx asc)
__global__ void __launch_bounds__(64) topk_merge_kernel(const float* __restrict__ ps, const long long* __restrict__ pi, int n_parts,
                                                        int nq, int k, float* __restrict__ out_s, long long* __restrict__ out_i) {
    __shared__ float cs[64 * kTopkMax];
    __shared__ long long ci[64 * kTopkMax];
    __shared__ float sel_s[kTopkMax];
    __shared__ long long sel_i[kTopkMax];
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

// Query operand of the fused scan: 256 rows (nq real ones, then zeros) x 512 as fp16, or e4m3 scaled by kF8Scale (stochastically
// rounded, see kF8LogP), row-major, plus per query the scan margin (accumulator units) and the certificate gap (cosine units; +inf on
// the fp16 copy, whose margin 2 eps |q| gmax is a deterministic bound and needs no certificate). One warp per row; runs once per search,
// so that the scan's CTAs fetch their operand with TMA instead of each converting the fp32 queries themselves.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// uniform 24-bit variates for the dither: pair p of row key `key` -> (r24 for the even element, r24 for the odd element)
__host__ __device__ __forceinline__ uint64_t dither_key(uint64_t seed, uint64_t row) {
    uint64_t z = seed ^ (row * 0xD6E8FEB86659FD93ull);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t dither_bits(uint64_t key, uint32_t pair) {
    uint64_t z = key + pair;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
constexpr uint64_t kQueryDitherSalt = 0x51ED270B7F4A7C15ull;  // queries and rows draw from different streams of the same seed

// stochastic e4m3 image of 16 elements of a 512-vector held as 4 float4 (element (lane + 32 i) * 4 + e), written to dst8 (row base);
// accumulates sum u^4, sum abar^4, sum abar^2 (scaled units) and whether anything saturated
__device__ __forceinline__ void f8_dither_row(const float4 (&a)[4], int lane, uint64_t key, uint8_t* __restrict__ dst8, float& u4,
                                              float& a4, float& a2, bool& sat) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t pair0 = static_cast<uint32_t>(lane + 32 * i) * 2u;
        const uint64_t h0 = dither_bits(key, pair0), h1 = dither_bits(key, pair0 + 1u);
        const float x[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
        const uint32_t r[4] = {static_cast<uint32_t>(h0) & 0xFFFFFFu, static_cast<uint32_t>(h0 >> 32) & 0xFFFFFFu,
                               static_cast<uint32_t>(h1) & 0xFFFFFFu, static_cast<uint32_t>(h1 >> 32) & 0xFFFFFFu};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float u, abar;
            v[e] = f8_round_dither(x[e], r[e], u, abar);
            sat |= fabsf(x[e]) * kF8Scale > kF8Max;
            const float uu = u * u, aa = abar * abar;
            u4 = fmaf(uu, uu, u4);
            a4 = fmaf(aa, aa, a4);
            a2 += aa;
        }
        // the values are exactly representable: the conversion below cannot round
        const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(v[0], v[1]), __NV_SATFINITE, __NV_E4M3);
        const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(v[2], v[3]), __NV_SATFINITE, __NV_E4M3);
        reinterpret_cast<uint32_t*>(dst8)[lane + 32 * i] = lo | (hi << 16);
    }
}

template <bool F8>
__global__ void __launch_bounds__(256) prep_queries_kernel(const float* __restrict__ q, int nq, const float* __restrict__ gmax_ptr,
                                                           const float* __restrict__ g4max_ptr, const float* __restrict__ w4max_ptr,
                                                           float scale, unsigned long long seed, void* __restrict__ q_img,
                                                           float* __restrict__ q_margin, float* __restrict__ q_gap,
                                                           int* __restrict__ flagged_acc_reset) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (flagged_acc_reset && blockIdx.x == 0 && threadIdx.x == 0) *flagged_acc_reset = 0;  // first chunk of a call
    if (r >= 2 * kQRows) return;
    float4 a[4];
    if (r < nq) load512(q + static_cast<size_t>(r) * kDim, lane, a);
    else {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (F8) {
        // scale = ln(1/p) (kF8LogP or its override). All norms below in cosine units (scaled sums / kF8Scale^n).
        float u4 = 0.f, a4 = 0.f, a2 = 0.f;
        bool sat = false;
        f8_dither_row(a, lane, dither_key(seed ^ kQueryDitherSalt, static_cast<uint64_t>(r)), static_cast<uint8_t*>(q_img) + static_cast<size_t>(r) * kDim,
                      u4, a4, a2, sat);
        u4 = warp_sum(u4);
        a4 = warp_sum(a4);
        a2 = warp_sum(a2);
        sat = __any_sync(0xffffffffu, sat);
        constexpr float kS2 = 1.f / (kF8Scale * kF8Scale), kS4 = kS2 * kS2;
        const float qbar4 = a4 * kS4, uq4 = u4 * kS4, qbar2 = a2 * kS2;
        // V <= |qbar|_4^2 W^2 + G4^2 |u(q)|_4^2 ; E = sqrt(V ln(1/p) / 2) + eps_det, rounded up
        const float V = sqrtf(qbar4 * __ldg(w4max_ptr)) + sqrtf(__ldg(g4max_ptr) * uq4);
        const float E = (sqrtf(0.5f * scale * V) + kF8AccEps * sqrtf(qbar2) * 1.125f * __ldg(gmax_ptr)) * 1.001f + 1e-7f;
        if (lane == 0) {
            q_margin[r] = r < nq ? (1.f + kF8GapFrac) * E * (kF8Scale * kF8Scale) : 0.f;
            // a saturated query component breaks the error model: the certificate can never pass and the exact scan answers
            q_gap[r] = sat ? -INFINITY : kF8GapFrac * E;
        }
    } else {  // scale = kCoarseEps: provable margin 2 eps |q| gmax
        const float qn = sqrtf(dot512(a, a));
        const float m = 2.f * scale * qn * __ldg(gmax_ptr);
        // the bound assumes fp16's NORMAL range: a query component beyond 65504 (inf in the operand image) or a query so small that its
        // components are fp16 subnormals (absolute error 2^-25 each instead of relative 2^-11) voids it -> always recomputed exactly
        float cm = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) cm = fmaxf(fmaxf(cm, fmaxf(fabsf(a[i].x), fabsf(a[i].y))), fmaxf(fabsf(a[i].z), fabsf(a[i].w)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
        const bool ok = cm <= 65504.f && (qn >= kF16MinNorm || qn == 0.f) && qn == qn;
        if (lane == 0) {
            q_margin[r] = r < nq ? m : 0.f;
            q_gap[r] = (ok || r >= nq) ? INFINITY : -INFINITY;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half2 lo = __floats2half2_rn(a[i].x, a[i].y), hi = __floats2half2_rn(a[i].z, a[i].w);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&lo);
            o.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(static_cast<__half*>(q_img) + static_cast<size_t>(r) * kDim)[lane + 32 * i] = o;
        }
    }
}

// e4m3 scan copy (rows x kF8Scale, stochastically rounded with the dither stream of (seed, global row id)). One warp per row.
// Also the bounds the certificate scales with, as the bits of non-negative floats (running maxima): g4max = largest sum of fourth
// powers of a row, w4max = largest sum of fourth powers of a row's e4m3 steps (both in cosine units).
__global__ void __launch_bounds__(256) make_f8_copy_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, long long n,
                                                           unsigned long long seed, long long first_row_id, float* __restrict__ g4max,
                                                           float* __restrict__ w4max) {
    const int lane = threadIdx.x & 31;
    const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    float g4 = 0.f, w4 = 0.f;
    for (long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps) {
        float4 a[4];
        load512(src + static_cast<size_t>(r) * kDim, lane, a);
        float4 a2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a2[i] = make_float4(a[i].x * a[i].x, a[i].y * a[i].y, a[i].z * a[i].z, a[i].w * a[i].w);
        g4 = fmaxf(g4, dot512(a2, a2));
        float u4 = 0.f, b4 = 0.f, b2 = 0.f;
        bool sat = false;
        f8_dither_row(a, lane, dither_key(seed, static_cast<uint64_t>(first_row_id + r)), dst + static_cast<size_t>(r) * kDim, u4, b4, b2, sat);
        constexpr float kS4 = 1.f / (kF8Scale * kF8Scale * kF8Scale * kF8Scale);
        w4 = fmaxf(w4, warp_sum(u4) * kS4);
    }
    if (lane == 0 && g4 > 0.f) atomicMax(reinterpret_cast<int*>(g4max), __float_as_int(g4 * 1.0001f));
    if (lane == 0 && w4 > 0.f) atomicMax(reinterpret_cast<int*>(w4max), __float_as_int(w4 * 1.0001f));
}

// scan copy + largest row norm (the margin of the coarse pass scales with it). One warp per row.
// amax: largest |component| (the fp16 copy is only valid while it stays below 65504, checked by the host side).
__global__ void __launch_bounds__(256) make_scan_copy_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n,
                                                             float* __restrict__ gmax, float* __restrict__ amax) {
    const int lane = threadIdx.x & 31;
    const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    float wmax = 0.f, cmax = 0.f;
    for (long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps) {
        float4 a[4];
        load512(src + static_cast<size_t>(r) * kDim, lane, a);
        if (dst) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __half2 lo = __floats2half2_rn(a[i].x, a[i].y), hi = __floats2half2_rn(a[i].z, a[i].w);
                uint2 o;
                o.x = *reinterpret_cast<uint32_t*>(&lo);
                o.y = *reinterpret_cast<uint32_t*>(&hi);
                reinterpret_cast<uint2*>(dst + static_cast<size_t>(r) * kDim)[lane + 32 * i] = o;
            }
        }
        wmax = fmaxf(wmax, dot512(a, a));
#pragma unroll
        for (int i = 0; i < 4; ++i) cmax = fmaxf(fmaxf(cmax, fmaxf(fabsf(a[i].x), fabsf(a[i].y))), fmaxf(fabsf(a[i].z), fabsf(a[i].w)));
    }
    // non-negative floats order like their bit patterns (a NaN / inf component ends up above every finite value: refused by the host)
    if (lane == 0 && wmax > 0.f) atomicMax(reinterpret_cast<int*>(gmax), __float_as_int(sqrtf(wmax) * 1.0000002f));
    if (amax && !(cmax == 0.f)) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(fabsf(cmax)));
}

}  // namespace frb
