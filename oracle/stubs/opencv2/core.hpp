// ORACLE BUILD ONLY. Stand-in for <opencv2/core.hpp> so that /root/reference/src/retinaface.cpp compiles verbatim on an image
// without OpenCV's C++ headers. Only RetinaFace::preprocess touches cv:: (resize / copyTo / convertTo / split); the oracle never
// calls preprocess (the detector's letterbox is pinned against the cv2 Python wheel instead), so these members do the
// simplest sensible thing. postprocessing / create_anchor_retinaface / nms — the functions the oracle drives — use only the
// MIN / MAX macros from here.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#define CV_8UC3 16
#define CV_32F 5

namespace cv {
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
};
struct Size {
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
class Mat {
  public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_(new std::vector<float>(static_cast<size_t>(r) * c * 3)) {}
    Mat(int r, int c, int type, const Scalar &s) : Mat(r, c, type) {
        for (size_t i = 0; i < buf_->size(); ++i) (*buf_)[i] = static_cast<float>(s.val[i % 3]);
    }
    Size size() const { return Size(cols, rows); }
    void release() { *this = Mat(); }
    Mat operator()(const Rect &) const { return *this; }
    void copyTo(Mat) const {}
    void convertTo(Mat &, int) const {}
    void push_back(const Mat &) {}
    template <class T>
    T *ptr(int = 0) { return buf_ ? reinterpret_cast<T *>(buf_->data()) : nullptr; }

  private:
    int type_ = CV_8UC3;
    std::shared_ptr<std::vector<float>> buf_;
};
inline Mat operator-(const Mat &m, const Scalar &) { return m; }
inline void split(const Mat &m, std::vector<Mat> &out) { out.assign(3, m); }
}  // namespace cv
