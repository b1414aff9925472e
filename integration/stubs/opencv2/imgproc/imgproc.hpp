// Minimal stand-in for <opencv2/imgproc/imgproc.hpp>: cvtColor(GRAY2BGR) only.
#ifndef SIVO_STUB_OPENCV_IMGPROC_HPP
#define SIVO_STUB_OPENCV_IMGPROC_HPP
#include "opencv2/core/core.hpp"
namespace cv {
enum { COLOR_GRAY2BGR = 8 };
inline void cvtColor(const Mat &src, Mat &dst, int code) {
    CV_Assert(code == COLOR_GRAY2BGR && src.type() == CV_8UC1);
    Mat out(src.rows, src.cols, CV_8UC3);
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) { const uint8_t v = src.at<uint8_t>(y, x); out.at<Vec3b>(y, x) = Vec3b(v, v, v); }
    dst = out;
}
}  // namespace cv
#endif
