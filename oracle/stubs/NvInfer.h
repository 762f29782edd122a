// Minimal stand-in for TensorRT's NvInfer.h so that the reference's src/common.h (which derives TRTLogger from
// nvinfer1::ILogger, /root/reference/src/common.h:28-53) compiles on a host without TensorRT. ORACLE BUILD ONLY.
#pragma once
#include <cstdint>
#include <cuda_runtime_api.h>
namespace nvinfer1 {
class ILogger {
  public:
    enum class Severity : int32_t { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
    virtual void log(Severity severity, const char *msg) noexcept = 0;
    virtual ~ILogger() = default;
};
}  // namespace nvinfer1
