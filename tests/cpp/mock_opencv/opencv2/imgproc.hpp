// TEST-ONLY stand-in for <opencv2/imgproc.hpp>: nearest-neighbour resize is enough for the compile/behaviour test of the shim.
#pragma once
#include "core.hpp"
namespace cv {
enum { INTER_LINEAR = 1, INTER_CUBIC = 2 };
inline void resize(const Mat &src, Mat &dst, Size sz, double = 0, double = 0, int = INTER_LINEAR) {
    Mat out(sz.height, sz.width, src.type());
    for (int r = 0; r < sz.height; ++r)
        for (int c = 0; c < sz.width; ++c)
            std::memcpy(out.data + out.step * r + c * out.elemSize(),
                        src.data + src.step * (r * src.rows / sz.height) + (c * src.cols / sz.width) * src.elemSize(), out.elemSize());
    dst = out;
}
}  // namespace cv
