// Minimal stand-ins for the three OpenCV types on the boundary (cv::Mat for 8-bit single-channel images and
// N x 32 descriptor matrices, cv::KeyPoint, cv::Point2f) so that include/gvs_adapter.hpp can be compiled and
// tested where C++ OpenCV is not installed.  With OpenCV present, define GVS_HAVE_OPENCV and this file is unused.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() = default; Point2f(float a, float b) : x(a), y(b) {} };

struct KeyPoint {  // same 28-byte layout as the real cv::KeyPoint
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};

enum { CV_8U = 0 };

class Mat {
  public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t* data = nullptr;
    Mat() = default;
    Mat(int r, int c, int /*type*/) { create(r, c); }
    void create(int r, int c) {
        rows = r; cols = c; step = (size_t)c;
        buf_ = std::shared_ptr<uint8_t>(new uint8_t[(size_t)r * c + 1], std::default_delete<uint8_t[]>());
        data = buf_.get();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    uint8_t* ptr(int r = 0) { return data + (size_t)r * step; }
    const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step; }
    Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)r * step; m.buf_ = buf_; return m; }
    Mat clone() const { Mat m(rows, cols, CV_8U); for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), cols); return m; }
  private:
    std::shared_ptr<uint8_t> buf_;
};

}  // namespace cv
