// oracle/shim/opencv2/core.hpp — TEST INFRASTRUCTURE ONLY (never linked into libgvs.so).
//
// Stand-in for the small part of the OpenCV C++ API that the reference's src/FeatureExtractor.cpp and
// src/FeatureMatcher.cpp use, so that those two files can be compiled UNMODIFIED, from where they lie under
// /root/reference, into oracle/_ref/ (C++ OpenCV is not installed in this image; see oracle/Makefile target `ref`).
// The container types (Mat, KeyPoint, Point_, Size, Rect) are written from the OpenCV documentation.  The five image
// primitives the path calls — resize(INTER_LINEAR), FAST(nms), GaussianBlur(7x7, sigma 2), sortIdx(DESCENDING),
// fastAtan2 — are dispatched through a hook table (oracle/ref_hooks.h): tests install callbacks into the real OpenCV
// (python cv2) to pin the oracle against "reference source + real OpenCV"; without cv2 the hooks default to the
// restatements in oracle/cpu_ref.cpp, themselves pinned against cv2 by tests/test_oracle_vs_cv2.py.
#pragma once
// (the real opencv2/core.hpp pulls in most of the standard library; the reference relies on that)
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../ref_hooks.h"

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795

// cvRound: round half to even (lrint in the default rounding mode), as OpenCV does with SSE2 / lrint.
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvFloor(float v) { return (int)std::floor(v); }
inline int cvFloor(int v) { return v; }
inline int cvCeil(double v) { return (int)std::ceil(v); }
inline int cvCeil(float v) { return (int)std::ceil(v); }
inline int cvCeil(int v) { return v; }

#define GVS_SHIM_REQUIRE(cond, what)                                                            \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            std::fprintf(stderr, "opencv shim: unsupported use: %s (%s:%d)\n", what, __FILE__, __LINE__); \
            std::abort();                                                                       \
        }                                                                                       \
    } while (0)

namespace cv {

template <class T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U>
    Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(double s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
    Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
template <class T> inline Point_<T> operator+(Point_<T> a, const Point_<T>& b) { return a += b; }
template <class T> inline Point_<T> operator-(Point_<T> a, const Point_<T>& b) { return a -= b; }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <class T>
struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
};

struct KeyPoint {
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 0;
};

enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
                   BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum SortFlags { SORT_EVERY_ROW = 0, SORT_EVERY_COLUMN = 1, SORT_ASCENDING = 0, SORT_DESCENDING = 16 };
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };

// Reference-counted 2-D array with ROI views (the subset of cv::Mat the path uses).
class Mat {
public:
    int flags = 0;   // only the type bits
    int rows = 0, cols = 0;
    size_t step = 0;  // bytes per row
    uchar* data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0)
        : flags(type), rows(r), cols(c), step(step_ ? step_ : (size_t)c * elemSizeOf(type)), data((uchar*)ext) {}

    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        for (int y = 0; y < r; ++y) std::memset(m.ptr(y), 0, (size_t)c * m.elemSize());
        return m;
    }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static Mat eye(int r, int c, int type) {
        GVS_SHIM_REQUIRE(type == CV_64F, "Mat::eye: only CV_64F");
        Mat m = zeros(r, c, type);
        for (int i = 0; i < std::min(r, c); ++i) m.at<double>(i, i) = 1.0;
        return m;
    }

    static size_t elemSizeOf(int type) {
        static const int depthBytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        return (size_t)depthBytes[type & 7] * (size_t)(((type >> CV_CN_SHIFT) & 511) + 1);
    }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return ((flags >> CV_CN_SHIFT) & 511) + 1; }
    size_t elemSize() const { return elemSizeOf(flags); }
    size_t elemSize1() const { return elemSizeOf(flags & 7); }
    size_t step1() const { return step / elemSize1(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }

    // Like cv::Mat::create: keeps the current buffer (and therefore a ROI view into a parent) if shape and type match.
    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && flags == type) return;
        flags = type; rows = r; cols = c;
        step = (size_t)c * elemSizeOf(type);
        size_t bytes = step * (size_t)r;
        owner_.reset(bytes ? new uchar[bytes] : nullptr, std::default_delete<uchar[]>());
        data = owner_.get();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { owner_.reset(); data = nullptr; rows = cols = 0; step = 0; }

    // cv::Mat::push_back(const Mat&): appends the rows of m (same type and width; an empty matrix adopts them)
    void push_back(const Mat& m) {
        if (m.empty()) return;
        GVS_SHIM_REQUIRE(empty() || (m.cols == cols && m.flags == flags), "push_back: type / width mismatch");
        Mat out(rows + m.rows, m.cols, m.flags);
        const size_t rb = (size_t)m.cols * m.elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(out.ptr(y), ptr(y), rb);
        for (int y = 0; y < m.rows; ++y) std::memcpy(out.ptr(rows + y), m.ptr(y), rb);
        *this = out;
    }

    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }

    Mat operator()(const Rect& roi) const {
        Mat m;
        m.flags = flags; m.rows = roi.height; m.cols = roi.width; m.step = step;
        m.data = data + (size_t)roi.y * step + (size_t)roi.x * elemSize();
        m.owner_ = owner_;
        return m;
    }
    // cv::Mat::rowRange / colRange assert 0 <= start <= end <= size; so does the stand-in (a reference that would throw aborts).
    Mat rowRange(int r0, int r1) const {
        GVS_SHIM_REQUIRE(0 <= r0 && r0 <= r1 && r1 <= rows, "rowRange out of range");
        return (*this)(Rect(0, r0, cols, r1 - r0));
    }
    Mat colRange(int c0, int c1) const {
        GVS_SHIM_REQUIRE(0 <= c0 && c0 <= c1 && c1 <= cols, "colRange out of range");
        return (*this)(Rect(c0, 0, c1 - c0, rows));
    }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }

    Mat clone() const {
        Mat m;
        copyTo(m);
        return m;
    }
    // OutputArray semantics: temporaries (e.g. dst.row(i)) are accepted and written through.
    void copyTo(const Mat& dst_) const {
        Mat& dst = const_cast<Mat&>(dst_);
        dst.create(rows, cols, flags);
        size_t bytes = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memmove(dst.ptr(y), ptr(y), bytes);
    }

private:
    std::shared_ptr<uchar> owner_;
};

inline int borderReflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// cv::copyMakeBorder for BORDER_REFLECT_101 (with or without BORDER_ISOLATED: the reference's sources are either whole
// images or flagged isolated).  src may be the interior view of dst, as in FeatureExtractor::computePyramid.
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType,
                           const Scalar& = Scalar()) {
    GVS_SHIM_REQUIRE((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101, "copyMakeBorder border type");
    const int es = (int)src.elemSize();
    dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    for (int y = 0; y < src.rows; ++y) {
        uchar* d = dst.ptr(y + top);
        if (d + (size_t)left * es != src.ptr(y)) std::memmove(d + (size_t)left * es, src.ptr(y), (size_t)src.cols * es);
        for (int x = 0; x < left; ++x)
            std::memcpy(d + (size_t)x * es, d + (size_t)(left + borderReflect101(x - left, src.cols)) * es, es);
        for (int x = 0; x < right; ++x)
            std::memcpy(d + (size_t)(left + src.cols + x) * es,
                        d + (size_t)(left + borderReflect101(src.cols + x, src.cols)) * es, es);
    }
    const size_t rowBytes = (size_t)dst.cols * es;
    for (int y = 0; y < top; ++y) std::memcpy(dst.ptr(y), dst.ptr(top + borderReflect101(y - top, src.rows)), rowBytes);
    for (int y = 0; y < bottom; ++y)
        std::memcpy(dst.ptr(top + src.rows + y), dst.ptr(top + borderReflect101(src.rows + y, src.rows)), rowBytes);
}

namespace shim_detail {
inline std::vector<uchar> packed(const Mat& m) {
    std::vector<uchar> v((size_t)m.rows * m.cols);
    for (int y = 0; y < m.rows; ++y) std::memcpy(v.data() + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
    return v;
}
inline void unpack(const std::vector<uchar>& v, Mat& m) {
    for (int y = 0; y < m.rows; ++y) std::memcpy(m.ptr(y), v.data() + (size_t)y * m.cols, (size_t)m.cols);
}
}  // namespace shim_detail

inline void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
    GVS_SHIM_REQUIRE(src.type() == CV_8UC1 && interpolation == INTER_LINEAR && fx == 0 && fy == 0 && dsize.width > 0,
                     "resize: only 8UC1 INTER_LINEAR with an explicit size");
    dst.create(dsize, src.type());
    std::vector<uchar> s = shim_detail::packed(src), d((size_t)dsize.width * dsize.height);
    gvs_ref_get_hooks()->resize_linear(s.data(), src.cols, src.rows, d.data(), dsize.width, dsize.height);
    shim_detail::unpack(d, dst);
}

inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0,
                         int borderType = BORDER_DEFAULT) {
    GVS_SHIM_REQUIRE(src.type() == CV_8UC1 && ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 &&
                         borderType == BORDER_REFLECT_101,
                     "GaussianBlur: only 8UC1 7x7 sigma 2 REFLECT_101");
    std::vector<uchar> s = shim_detail::packed(src), d(s.size());
    gvs_ref_get_hooks()->gaussian_blur7(s.data(), src.cols, src.rows, d.data());
    dst.create(src.rows, src.cols, src.type());
    shim_detail::unpack(d, dst);
}

inline void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    GVS_SHIM_REQUIRE(image.type() == CV_8UC1 && nonmaxSuppression, "FAST: only 8UC1 with non-max suppression");
    keypoints.clear();
    if (image.rows < 7 || image.cols < 7) return;
    std::vector<uchar> s = shim_detail::packed(image);
    const int cap = image.rows * image.cols;
    std::vector<int> xs(cap), ys(cap), sc(cap);
    int n = gvs_ref_get_hooks()->fast_nms(s.data(), image.cols, image.rows, threshold, xs.data(), ys.data(), sc.data(), cap);
    keypoints.reserve(n);
    for (int i = 0; i < n; ++i) keypoints.emplace_back((float)xs[i], (float)ys[i], 7.f, -1.f, (float)sc[i]);
}

inline void sortIdx(const std::vector<float>& src, std::vector<int>& dst, int flags) {
    GVS_SHIM_REQUIRE(flags == SORT_DESCENDING || flags == (SORT_DESCENDING | SORT_EVERY_ROW), "sortIdx: only DESCENDING");
    dst.resize(src.size());
    if (!src.empty()) gvs_ref_get_hooks()->sort_idx_desc(src.data(), (int)src.size(), dst.data());
}

inline float fastAtan2(float y, float x) { return gvs_ref_get_hooks()->fast_atan2(y, x); }

// cv::norm(a, b, NORM_L1) on 8-bit windows: an exact integer sum of absolute differences, returned as double.
inline double norm(const Mat& a, const Mat& b, int normType) {
    GVS_SHIM_REQUIRE(normType == NORM_L1 && a.type() == CV_8UC1 && b.type() == CV_8UC1 && a.rows == b.rows && a.cols == b.cols,
                     "norm: only NORM_L1 on equal-size 8UC1");
    long s = 0;
    for (int y = 0; y < a.rows; ++y) {
        const uchar* pa = a.ptr(y); const uchar* pb = b.ptr(y);
        for (int x = 0; x < a.cols; ++x) s += std::abs((int)pa[x] - (int)pb[x]);
    }
    return (double)s;
}

// Declared so that the reference's Conversions.h parses; nothing on the matcher path executes them.
inline void Rodrigues(const Mat&, Mat&) { GVS_SHIM_REQUIRE(false, "Rodrigues"); }
template <class E> inline void cv2eigen(const Mat&, E&) { GVS_SHIM_REQUIRE(false, "cv2eigen"); }
template <class E> inline void eigen2cv(const E&, Mat&) { GVS_SHIM_REQUIRE(false, "eigen2cv"); }

}  // namespace cv
