// TEST-ONLY stand-in for <opencv2/core.hpp>: this image has no OpenCV C++ headers (SURVEY §7 hard part 8), so the drop-in
// compile test uses this minimal cv::Mat (continuous or ROI view, u8 / f32) — just what the shim headers and an app.cpp-style
// caller touch. It is NOT part of the product; a real build includes the real OpenCV.
#pragma once
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8UC3 16
#define CV_32F 5
#define CV_32FC1 5

namespace cv {
struct Point {
    int x, y;
    Point(int x_ = 0, int y_ = 0) : x(x_), y(y_) {}
};
struct Size {
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(Point a, Point b) : x(a.x < b.x ? a.x : b.x), y(a.y < b.y ? a.y : b.y), width(std::abs(a.x - b.x)), height(std::abs(a.y - b.y)) {}
};
class Mat {
  public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        rows = r;
        cols = c;
        type_ = type;
        step = static_cast<size_t>(c) * elemSize();
        buf_ = std::shared_ptr<std::vector<unsigned char>>(new std::vector<unsigned char>(step * r));
        data = buf_->data();
    }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_8UC3 ? 3 : 4; }
    bool isContinuous() const { return step == static_cast<size_t>(cols) * elemSize(); }
    bool empty() const { return data == nullptr; }
    template <class T>
    T *ptr(int r = 0) { return reinterpret_cast<T *>(data + step * r); }
    template <class T>
    const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + step * r); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + m.step * r, data + step * r, static_cast<size_t>(cols) * elemSize());
        return m;
    }
    Mat operator()(const Rect &roi) const {
        Mat m = *this;
        m.rows = roi.height;
        m.cols = roi.width;
        m.data = data + step * roi.y + roi.x * elemSize();
        return m;
    }
    void push_back(const Mat &o) {  // append rows
        Mat m(rows + o.rows, o.cols, o.type_);
        if (rows) std::memcpy(m.data, data, step * rows);
        for (int r = 0; r < o.rows; ++r) std::memcpy(m.data + m.step * (rows + r), o.data + o.step * r, m.step);
        *this = m;
    }
    void release() { *this = Mat(); }

  private:
    int type_ = CV_8UC3;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};
}  // namespace cv
