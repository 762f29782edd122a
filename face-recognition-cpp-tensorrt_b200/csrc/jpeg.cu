// JPEG decode for the serving loop: the reference decodes every request on the CPU with cv::imdecode and stretches it
// to the configured frame size with cv::resize (/root/reference src/app.cpp:247-256 "/recognize", :294-301 "/inference"). Here the
// Huffman stage runs on the host and the IDCT / colour conversion on the GPU through nvJPEG (a CUDA toolkit library, loaded with dlopen
// at first use so that libfr_b200.so itself has no link-time dependency on it); the optional resize is the library's own kernel with
// OpenCV's u8 INTER_LINEAR fixed-point arithmetic (det_letterbox_kernel, bit-exact against cv2.resize when shrinking), so that
//   fr_jpeg_decode(dec, jpeg, n, W, H, frame, stride)  ==  cv::resize(cv::imdecode(jpeg), Size(W, H))   up to the JPEG decoder's
// own IDCT / chroma-upsampling rounding (libjpeg-turbo and nvJPEG are different third-party decoders; the test states the tolerance).
#include <dlfcn.h>
#include <nvjpeg.h>

#include <memory>
#include <string>

#include "common.h"

using namespace frb;

namespace {

struct NvJpegApi {
    void* lib = nullptr;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                             cudaStream_t) = nullptr;
};

// one process-wide binding; the library stays loaded
const NvJpegApi& nvjpeg_api() {
    static const NvJpegApi api = [] {
        NvJpegApi a;
        for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib) throw FileError{FR_ENOENT, "JPEG decode needs nvJPEG (libnvjpeg.so.12 of the CUDA toolkit): not found"};
        auto sym = [&](const char* n) {
            void* p = dlsym(a.lib, n);
            if (!p) throw FileError{FR_ENOENT, std::string("nvJPEG symbol missing: ") + n};
            return p;
        };
        a.CreateSimple = reinterpret_cast<decltype(a.CreateSimple)>(sym("nvjpegCreateSimple"));
        a.Destroy = reinterpret_cast<decltype(a.Destroy)>(sym("nvjpegDestroy"));
        a.JpegStateCreate = reinterpret_cast<decltype(a.JpegStateCreate)>(sym("nvjpegJpegStateCreate"));
        a.JpegStateDestroy = reinterpret_cast<decltype(a.JpegStateDestroy)>(sym("nvjpegJpegStateDestroy"));
        a.GetImageInfo = reinterpret_cast<decltype(a.GetImageInfo)>(sym("nvjpegGetImageInfo"));
        a.Decode = reinterpret_cast<decltype(a.Decode)>(sym("nvjpegDecode"));
        return a;
    }();
    return api;
}

void check_nvjpeg(nvjpegStatus_t s, const char* what) {
    if (s == NVJPEG_STATUS_SUCCESS) return;
    // a stream that is not a decodable JPEG is the caller's input error: the reference throws "Empty image" (src/app.cpp:251,298)
    if (s == NVJPEG_STATUS_BAD_JPEG || s == NVJPEG_STATUS_JPEG_NOT_SUPPORTED || s == NVJPEG_STATUS_INVALID_PARAMETER)
        throw ArgError{std::string("Empty image (") + what + ": not a decodable JPEG, nvjpeg status " + std::to_string(static_cast<int>(s)) + ")"};
    throw CudaError{std::string(what) + " failed: nvjpeg status " + std::to_string(static_cast<int>(s))};
}

}  // namespace

struct FrJpegDecoder {
    int device = 0;
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    cudaStream_t stream = nullptr;
    uint8_t* raw = nullptr;  // decoded image, BGR interleaved, w x h
    size_t raw_cap = 0;
    uint8_t* out = nullptr;  // resized image
    size_t out_cap = 0;
};

namespace {
void reserve(uint8_t*& p, size_t& cap, size_t need) {
    if (cap >= need) return;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    FRB_CUDA(cudaMalloc(&p, need));
    cap = need;
}
}  // namespace

extern "C" {

int fr_jpeg_decoder_create(int device, FrJpegDecoder** out) {
    return guarded([&] {
        if (!out) throw ArgError{"out is null"};
        use_device(device);  // sm_100 only, like every handle of the library
        const NvJpegApi& api = nvjpeg_api();
        std::unique_ptr<FrJpegDecoder> d(new FrJpegDecoder());
        d->device = device;
        DeviceGuard dg(device);
        try {
            check_nvjpeg(api.CreateSimple(&d->handle), "nvjpegCreateSimple");
            check_nvjpeg(api.JpegStateCreate(d->handle, &d->state), "nvjpegJpegStateCreate");
            FRB_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        } catch (...) {
            fr_jpeg_decoder_destroy(d.release());
            throw;
        }
        *out = d.release();
    });
}

void fr_jpeg_decoder_destroy(FrJpegDecoder* d) {
    if (!d) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    if (d->state) nvjpeg_api().JpegStateDestroy(d->state);
    if (d->handle) nvjpeg_api().Destroy(d->handle);
    if (d->stream) cudaStreamDestroy(d->stream);
    cudaFree(d->raw);
    cudaFree(d->out);
    if (prev >= 0) cudaSetDevice(prev);
    delete d;
}

int fr_jpeg_info(FrJpegDecoder* d, const uint8_t* jpeg, size_t nbytes, int* width, int* height) {
    return guarded([&] {
        if (!d || !width || !height) throw ArgError{"null argument"};
        if (!jpeg || nbytes == 0) throw ArgError{"Empty image"};
        int ncomp = 0, w[NVJPEG_MAX_COMPONENT] = {}, h[NVJPEG_MAX_COMPONENT] = {};
        nvjpegChromaSubsampling_t ss;
        check_nvjpeg(nvjpeg_api().GetImageInfo(d->handle, jpeg, nbytes, &ncomp, &ss, w, h), "nvjpegGetImageInfo");
        if (w[0] <= 0 || h[0] <= 0) throw ArgError{"Empty image"};
        *width = w[0];
        *height = h[0];
    });
}

int fr_jpeg_decode(FrJpegDecoder* d, const uint8_t* jpeg, size_t nbytes, int out_w, int out_h, uint8_t* bgr, int stride) {
    return guarded([&] {
        NvtxRange nvtx("fr.jpeg.decode");
        if (!d || !bgr) throw ArgError{"null argument"};
        if (!jpeg || nbytes == 0) throw ArgError{"Empty image"};
        const NvJpegApi& api = nvjpeg_api();
        DeviceGuard dg(d->device);
        int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {}, hs[NVJPEG_MAX_COMPONENT] = {};
        nvjpegChromaSubsampling_t ss;
        check_nvjpeg(api.GetImageInfo(d->handle, jpeg, nbytes, &ncomp, &ss, ws, hs), "nvjpegGetImageInfo");
        const int w = ws[0], h = hs[0];
        if (w <= 0 || h <= 0) throw ArgError{"Empty image"};
        if (out_w <= 0 || out_h <= 0) {  // "as decoded"
            out_w = w;
            out_h = h;
        }
        if (stride < out_w * 3) throw ArgError{"stride is smaller than a row of the output frame"};
        reserve(d->raw, d->raw_cap, static_cast<size_t>(w) * h * 3);
        nvjpegImage_t img{};
        img.channel[0] = d->raw;
        img.pitch[0] = static_cast<size_t>(w) * 3;
        check_nvjpeg(api.Decode(d->handle, d->state, jpeg, nbytes, NVJPEG_OUTPUT_BGRI, &img, d->stream), "nvjpegDecode");
        const uint8_t* src = d->raw;
        if (out_w != w || out_h != h) {
            // cv::resize(frame, frame, Size(W, H)) (src/app.cpp:255,301): INTER_LINEAR, stretch to the whole target (no letterbox)
            reserve(d->out, d->out_cap, static_cast<size_t>(out_w) * out_h * 3);
            launch_stretch_resize_u8(d->raw, h, w, w * 3, out_h, out_w, d->out, d->stream);
            src = d->out;
        }
        FRB_CUDA(cudaMemcpy2DAsync(bgr, stride, src, static_cast<size_t>(out_w) * 3, static_cast<size_t>(out_w) * 3, out_h, cudaMemcpyDeviceToHost,
                                   d->stream));
        FRB_CUDA(cudaStreamSynchronize(d->stream));
    });
}

}  // extern "C"
