// Read-only HBM bandwidth probe (calibration for the scan kernel's roofline): streams a buffer much larger than L2 with
// (a) 128-bit LDG from a grid-stride loop and (b) 16 KiB cp.async.bulk copies into a shared-memory ring (no compute), and
// prints GB/s for each. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/hbm_read tools/hbm_read_peak.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ p, size_t n, uint32_t* out) {
    uint32_t acc = 0;
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n; i += stride) {
        const uint4 a = __ldcs(p + i);
        acc ^= a.x ^ a.y ^ a.z ^ a.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// one CTA per SM, one thread issues 16 KiB bulk copies into a ring of `kStages` slots and waits for them in order
template <int kStages>
__global__ void __launch_bounds__(128) bulk_kernel(const uint8_t* __restrict__ p, size_t chunks, uint32_t* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[kStages];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        size_t issued = blockIdx.x, waited = blockIdx.x;
        int si = 0, sw = 0;
        uint32_t phase = 0;
        int inflight = 0;
        while (waited < chunks) {
            while (inflight < kStages && issued < chunks) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[si])), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(smem + si * 16384)),
                             "l"(p + issued * 16384), "r"(16384), "r"(smem_u32(&bar[si]))
                             : "memory");
                issued += gridDim.x;
                si = (si + 1) % kStages;
                ++inflight;
            }
            uint32_t ok = 0;
            while (!ok) {
                asm volatile(
                    "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                    : "=r"(ok)
                    : "r"(smem_u32(&bar[sw])), "r"(phase)
                    : "memory");
            }
            waited += gridDim.x;
            --inflight;
            if (++sw == kStages) {
                sw = 0;
                phase ^= 1;
            }
        }
        if (smem[5] == 77 && smem[16384 + 9] == 78) out[1] = 1;
    }
}

int main() {
    const size_t bytes = 8ull << 30;
    uint8_t* buf;
    uint32_t* out;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&out, 64);
    cudaMemset(buf, 1, bytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time = [&](auto&& fn, const char* name) {
        for (int i = 0; i < 3; ++i) fn();
        cudaEventRecord(e0);
        const int reps = 10;
        for (int i = 0; i < reps; ++i) fn();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"probe\": \"%s\", \"GBps\": %.1f, \"ms\": %.4f, \"err\": \"%s\"}\n", name, bytes / (ms / reps * 1e-3) / 1e9, ms / reps,
               cudaGetErrorString(cudaGetLastError()));
    };
    for (int mult : {4, 8, 16})
        time([&] { ldg_kernel<<<sms * mult, 512>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, out); },
             mult == 4 ? "ldg128 x4 ctas/sm" : (mult == 8 ? "ldg128 x8 ctas/sm" : "ldg128 x16 ctas/sm"));
    cudaFuncSetAttribute(bulk_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 16384);
    cudaFuncSetAttribute(bulk_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384);
    time([&] { bulk_kernel<6><<<sms, 128, 6 * 16384>>>(buf, bytes / 16384, out); }, "bulk 16KiB x6 in flight per SM");
    time([&] { bulk_kernel<12><<<sms, 128, 12 * 16384>>>(buf, bytes / 16384, out); }, "bulk 16KiB x12 in flight per SM");
    time([&] { cudaMemcpyAsync(buf, buf + bytes / 2, bytes / 2, cudaMemcpyDeviceToDevice); }, "cudaMemcpy D2D (bytes = read+write of 4 GiB)");
    return 0;
}
