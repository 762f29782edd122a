#include "common.h"

#include <mutex>

namespace frb {

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }

int use_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw FileError{FR_ENODEVICE, "no CUDA device visible: this library has no CPU fallback"};
    }
    if (device < 0 || device >= count) throw ArgError{"device index out of range"};
    cudaDeviceProp prop{};
    FRB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10 || prop.minor != 0) {  // sm_100a cubins do not run on other 10.x parts
        throw FileError{FR_ENODEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                          ", kernels are built for sm_100a only"};
    }
    FRB_CUDA(cudaSetDevice(device));
    return prop.multiProcessorCount;
}

// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so that the library
// does not need to link libcuda (absent on CPU-only build hosts).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    if (!fn) throw CudaError{"cuTensorMapEncodeTiled entry point not available"};
    return fn;
}

CUtensorMap make_tmap_2d_f16(const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError{"cuTensorMapEncodeTiled(2d) failed with CUresult " + std::to_string(static_cast<int>(r))};
    return m;
}

CUtensorMap make_tmap_2d_u8(const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError{"cuTensorMapEncodeTiled(2d u8) failed with CUresult " + std::to_string(static_cast<int>(r))};
    return m;
}

CUtensorMap make_tmap_nhwc_f16(const void* base, uint64_t n, uint64_t h, uint64_t w, uint64_t c, uint32_t bn, uint32_t bh, uint32_t bw,
                               uint32_t bc) {
    CUtensorMap m;
    cuuint64_t gdim[4] = {c, w, h, n};
    cuuint64_t gstride[3] = {c * 2, w * c * 2, h * w * c * 2};
    cuuint32_t box[4] = {bc, bw, bh, bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError{"cuTensorMapEncodeTiled(4d) failed with CUresult " + std::to_string(static_cast<int>(r))};
    return m;
}

}  // namespace frb

extern "C" {
const char* fr_last_error(void) { return frb::t_last_error.c_str(); }
int fr_abi_version(void) { return 2; }
uint64_t fr_launch_count(void) { return frb::g_launches.load(); }
}
