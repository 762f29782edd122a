// ORACLE BUILD ONLY. Stand-in for <opencv2/imgproc.hpp> (see core.hpp beside it): RetinaFace::preprocess is never run by the oracle.
#pragma once
#include "core.hpp"
namespace cv {
enum { INTER_LINEAR = 1, INTER_CUBIC = 2 };
inline void resize(const Mat &, Mat &, Size, double = 0, double = 0, int = INTER_LINEAR) {}
}  // namespace cv
