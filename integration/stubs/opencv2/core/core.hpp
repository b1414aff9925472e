// Minimal stand-in for <opencv2/core/core.hpp>: only what integration/src/*.cpp use (see ../../README.md).
#ifndef SIVO_STUB_OPENCV_CORE_HPP
#define SIVO_STUB_OPENCV_CORE_HPP
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error(std::string("CV_Assert failed: ") + #expr); } while (0)

namespace cv {

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size &o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size &o) const { return !(*this == o); }
};
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() = default;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Point2f { float x = 0, y = 0; };
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
};
struct Vec3b {
    uint8_t val[3];
    Vec3b() : val{0, 0, 0} {}
    Vec3b(uint8_t a, uint8_t b, uint8_t c) : val{a, b, c} {}
    uint8_t &operator[](int i) { return val[i]; }
    const uint8_t &operator[](int i) const { return val[i]; }
};
// cv::KeyPoint: pt (2 floats), size, angle, response, octave, class_id = 28 bytes (types.hpp)
struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

inline int cv_depth_bytes(int type) { return (type & 7) == CV_64F ? 8 : 1; }
inline int cv_channels(int type) { return (type >> CV_CN_SHIFT) + 1; }

class _OutputArray;

class Mat {
 public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t *data = nullptr;

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &s) {
        create(r, c, type);
        const int cn = cv_channels(type);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x)
                for (int k = 0; k < cn; ++k) {
                    if ((type & 7) == CV_64F) reinterpret_cast<double *>(data + y * step)[x * cn + k] = s.val[k];
                    else (data + y * step)[x * cn + k] = static_cast<uint8_t>(s.val[k]);
                }
    }
    // header over caller memory (no ownership), like cv::Mat(rows, cols, type, void*, step)
    Mat(int r, int c, int type, void *ext, size_t st = 0) : rows(r), cols(c), data(static_cast<uint8_t *>(ext)), type_(type) {
        step = st ? st : static_cast<size_t>(c) * elemSize();
    }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_ && owner_ && step == static_cast<size_t>(c) * elemSize()) return;
        rows = r; cols = c; type_ = type;
        step = static_cast<size_t>(c) * elemSize();
        owner_ = std::shared_ptr<uint8_t>(static_cast<uint8_t *>(std::malloc(std::max<size_t>(1, step * r))), std::free);
        data = owner_.get();
    }
    int type() const { return type_; }
    int channels() const { return cv_channels(type_); }
    size_t elemSize() const { return static_cast<size_t>(cv_depth_bytes(type_)) * cv_channels(type_); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return static_cast<size_t>(rows) * cols; }
    bool isContinuous() const { return step == static_cast<size_t>(cols) * elemSize(); }
    template <typename T> T &at(int i) { return rows == 1 ? reinterpret_cast<T *>(data)[i] : *reinterpret_cast<T *>(data + i * step); }
    template <typename T> T &at(int y, int x) { return reinterpret_cast<T *>(data + y * step)[x]; }
    template <typename T> const T &at(int y, int x) const { return reinterpret_cast<const T *>(data + y * step)[x]; }
    template <typename T> T *ptr(int y = 0) { return reinterpret_cast<T *>(data + y * step); }
    template <typename T> const T *ptr(int y = 0) const { return reinterpret_cast<const T *>(data + y * step); }
    Mat operator()(const Rect &r) const {  // ROI sharing the storage
        CV_Assert(r.x >= 0 && r.y >= 0 && r.x + r.width <= cols && r.y + r.height <= rows);
        Mat m;
        m.rows = r.height; m.cols = r.width; m.step = step; m.type_ = type_; m.owner_ = owner_;
        m.data = data + r.y * step + r.x * elemSize();
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat clone() const {
        Mat m;
        copyTo(m);
        return m;
    }
    void copyTo(Mat &dst) const {
        dst.create(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(dst.data + y * dst.step, data + y * step, static_cast<size_t>(cols) * elemSize());
    }
    void copyTo(const _OutputArray &dst) const;
    void release() { rows = cols = 0; step = 0; data = nullptr; owner_.reset(); }

 private:
    int type_ = 0;
    std::shared_ptr<uint8_t> owner_;
};

// Proxy arguments: the shim only needs Mat in, Mat out.
class _InputArray {
 public:
    _InputArray() = default;
    _InputArray(const Mat &m) : m_(&m) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
 private:
    const Mat *m_ = nullptr;
};
class _OutputArray {
 public:
    _OutputArray(Mat &m) : m_(&m) {}
    void release() const { m_->release(); }
    Mat &getMatRef() const { return *m_; }
 private:
    Mat *m_;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
inline void Mat::copyTo(const _OutputArray &dst) const { copyTo(dst.getMatRef()); }

enum { NORM_MINMAX = 32 };
// y = (x - min) * (b - a) / (max - min) + a  (cv::normalize with NORM_MINMAX on a single-channel double matrix)
inline void normalize(const Mat &src, Mat &dst, double a, double b, int norm_type, int dtype) {
    CV_Assert(norm_type == NORM_MINMAX && src.type() == CV_64FC1 && dtype == CV_64FC1);
    double lo = 1e300, hi = -1e300;
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) { lo = std::min(lo, src.at<double>(y, x)); hi = std::max(hi, src.at<double>(y, x)); }
    Mat out(src.rows, src.cols, CV_64FC1);
    const double sc = hi > lo ? (b - a) / (hi - lo) : 0.0;
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) out.at<double>(y, x) = (src.at<double>(y, x) - lo) * sc + a;
    dst = out;
}
// per-channel table look-up (cv::LUT with a 256x1 table of the same channel count)
inline void LUT(const Mat &src, const Mat &lut, Mat &dst) {
    CV_Assert(src.type() == CV_8UC3 && lut.type() == CV_8UC3 && lut.total() == 256);
    Mat out(src.rows, src.cols, CV_8UC3);
    const Vec3b *t = lut.ptr<Vec3b>();
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) {
            const Vec3b &s = src.at<Vec3b>(y, x);
            out.at<Vec3b>(y, x) = Vec3b(t[s[0]][0], t[s[1]][1], t[s[2]][2]);
        }
    dst = out;
}
inline uint8_t saturate_u8(double v) { return static_cast<uint8_t>(std::min(255.0, std::max(0.0, std::nearbyint(v)))); }
inline void addWeighted(const Mat &a, double alpha, const Mat &b, double beta, double gamma, Mat &dst) {
    CV_Assert(a.type() == CV_8UC3 && b.type() == CV_8UC3 && a.rows == b.rows && a.cols == b.cols);
    Mat out(a.rows, a.cols, CV_8UC3);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols * 3; ++x)
            out.ptr<uint8_t>(y)[x] = saturate_u8(a.ptr<uint8_t>(y)[x] * alpha + b.ptr<uint8_t>(y)[x] * beta + gamma);
    dst = out;
}
}  // namespace cv

#endif
