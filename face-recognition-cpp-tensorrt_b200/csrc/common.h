// Host-side helpers shared by the translation units of libfr_b200: error plumbing for the C ABI,
// launch counting, TMA descriptor construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <array>
#include <atomic>
#include <cstdlib>
#include <map>
#include <utility>
#include <cstdint>
#include <cstdio>
#include <string>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a profiler attaches (tracing)

#include "../../include/fr_b200.h"

namespace frb {

void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

// NVTX range over a stage of the hot path (host-side scope: what nsys / ncu --nvtx show as detect / crop / embed / search / exchange)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct CudaError {
    std::string msg;
};

#define FRB_CUDA(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) {                                                                             \
            throw ::frb::CudaError{std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                   std::to_string(__LINE__) + ")"};                                          \
        }                                                                                                    \
    } while (0)

struct ArgError {
    std::string msg;
};
struct StateError {
    std::string msg;
};
struct FileError {
    int code;
    std::string msg;
};

// Wrap the body of every extern "C" entry point: map C++ exceptions to ABI error codes.
template <class F>
int guarded(F&& f) {
    try {
        f();
        return FR_OK;
    } catch (const CudaError& e) {
        set_error(e.msg);
        cudaGetLastError();  // clear sticky-less errors
        return FR_ECUDA;
    } catch (const ArgError& e) {
        set_error(e.msg);
        return FR_EINVAL;
    } catch (const StateError& e) {
        set_error(e.msg);
        return FR_ESTATE;
    } catch (const FileError& e) {
        set_error(e.msg);
        return e.code;
    } catch (const std::exception& e) {
        set_error(std::string("internal: ") + e.what());
        return FR_ECUDA;
    }
}

// Launch-bound kernel sequences (the ~60-100 small launches of one network forward) are captured once per shape into a CUDA
// graph and replayed. `body` enqueues the kernels on `st`; keys identify everything baked into the launch arguments.
struct GraphCache {
    struct Entry {
        cudaGraphExec_t exec;
        int launches;
    };
    std::map<std::array<uint64_t, 3>, Entry> entries;
    bool enabled = std::getenv("FR_NO_GRAPHS") == nullptr;
    template <class F>
    void run(const std::array<uint64_t, 3>& key, cudaStream_t st, F&& body) {
        auto it = entries.find(key);
        if (!enabled || (it == entries.end() && entries.size() >= 16)) {
            body();
            return;
        }
        if (it == entries.end()) {
            const uint64_t before = g_launches.load();
            cudaGraph_t graph = nullptr;
            FRB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try {
                body();
            } catch (...) {
                cudaStreamEndCapture(st, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            FRB_CUDA(cudaStreamEndCapture(st, &graph));
            Entry e{};
            e.launches = static_cast<int>(g_launches.load() - before);
            g_launches.store(before);  // capture launched nothing; every replay accounts for its kernels below
            cudaError_t err = cudaGraphInstantiate(&e.exec, graph, 0);
            cudaGraphDestroy(graph);
            FRB_CUDA(err);
            it = entries.emplace(key, e).first;
        }
        FRB_CUDA(cudaGraphLaunch(it->second.exec, st));
        count_launch(it->second.launches);
    }
    ~GraphCache() {
        for (auto& kv : entries) cudaGraphExecDestroy(kv.second.exec);
    }
};

// Kernel launch with (pdl = true) programmatic dependent launch: the kernel may begin while its predecessor in the stream is still
// running and must call griddep_wait() (ptx_sm100.cuh) before it touches anything the predecessor reads or writes. Captured into CUDA
// graphs as programmatic dependency edges. OFF by default (FR_PDL=1 enables it): measured on the IR-SE-50 forward it gains 2 % at batch
// 32 (1.18 vs 1.20 ms) and LOSES 4-5 % at batch 256 (5.04 vs 4.85 ms through the host-buffer call, 4.59 vs 4.36 ms device-resident) -
// the persistent conv kernels fill every SM, so an early-launched successor has nowhere to run its prologue.
inline bool pdl_enabled() {
    static const bool on = std::getenv("FR_PDL") != nullptr && std::atoi(std::getenv("FR_PDL")) != 0;
    return on;
}
template <class... KArgs, class... Args>
void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (pdl && pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    FRB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...));
}

// Select `device`, check it is sm_100, return its SM count.
int use_device(int device);

// cv::resize(src, dst, Size(out_w, out_h)) for u8 BGR frames on the device (INTER_LINEAR, OpenCV's fixed-point arithmetic): the
// detector's letterbox kernel with the resized area covering the whole canvas (defined in detector.cu, used by jpeg.cu)
void launch_stretch_resize_u8(const uint8_t* src, int h, int w, int stride_bytes, int out_h, int out_w, uint8_t* dst, cudaStream_t st);

// RAII device switch for entry points
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) FRB_CUDA(cudaSetDevice(dev));
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// 2-D row-major 16-bit tensor [rows, cols] -> TMA map with box [box_rows, box_cols], 128-byte swizzle.
CUtensorMap make_tmap_2d_f16(const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);
// 2-D row-major byte tensor [rows, cols] -> TMA map with box [box_rows, box_cols <= 128], 128-byte swizzle (fp8 scan copy).
CUtensorMap make_tmap_2d_u8(const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);
// 4-D NHWC 16-bit tensor [n, h, w, c] -> TMA map with box [bn, bh, bw, bc], 128-byte swizzle, zero OOB fill.
CUtensorMap make_tmap_nhwc_f16(const void* base, uint64_t n, uint64_t h, uint64_t w, uint64_t c, uint32_t bn, uint32_t bh, uint32_t bw,
                               uint32_t bc);

}  // namespace frb
