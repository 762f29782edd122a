// Minimal stand-in for TensorRT's NvInfer.h so that the reference's own sources compile on a host without TensorRT:
//   src/common.h derives TRTLogger from nvinfer1::ILogger (/root/reference/src/common.h:28-53);
//   src/retinaface.cpp calls IRuntime / ICudaEngine / IExecutionContext in loadEngine, preInference and doInference
//   (/root/reference/src/retinaface.cpp:31-104,138-145).
// The classes below only have to make those call sites compile and let the constructor run to completion; no network is ever
// executed through them (the oracle drives postprocessing / create_anchor_retinaface / nms only). ORACLE BUILD ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cuda_runtime_api.h>
namespace nvinfer1 {
class ILogger {
  public:
    enum class Severity : int32_t { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
    virtual void log(Severity severity, const char *msg) noexcept = 0;
    virtual ~ILogger() = default;
};
class Dims4 {
  public:
    int d[4];
    Dims4(int a, int b, int c, int e) : d{a, b, c, e} {}
};
class IExecutionContext {
  public:
    bool enqueueV2(void **, cudaStream_t, cudaEvent_t *) noexcept { return false; }
    bool setBindingDimensions(int, Dims4) noexcept { return true; }
};
class ICudaEngine {
  public:
    IExecutionContext *createExecutionContext() noexcept { return &ctx_; }
    int getNbBindings() const noexcept { return nb_; }
    // binding order of the exported engines: input first, then the outputs in export order
    int getBindingIndex(const char *) noexcept { return next_ < nb_ ? next_++ : nb_ - 1; }
    void setNbBindings(int n) { nb_ = n; }

  private:
    IExecutionContext ctx_;
    int nb_ = 3, next_ = 0;
};
class IRuntime {
  public:
    ICudaEngine *deserializeCudaEngine(const void *, std::size_t) noexcept { return new ICudaEngine(); }
};
inline IRuntime *createInferRuntime(ILogger &) noexcept {
    static IRuntime rt;
    return &rt;
}
}  // namespace nvinfer1
