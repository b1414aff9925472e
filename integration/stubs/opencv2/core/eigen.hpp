// Minimal stand-in for <opencv2/core/eigen.hpp>: eigen2cv for row-major dynamic matrices.
#ifndef SIVO_STUB_OPENCV_EIGEN_HPP
#define SIVO_STUB_OPENCV_EIGEN_HPP
#include <Eigen/Core>
#include <cstdint>
#include "opencv2/core/core.hpp"
namespace cv {
template <typename T> struct stub_type;
template <> struct stub_type<double> { enum { value = CV_64FC1 }; };
template <> struct stub_type<uint8_t> { enum { value = CV_8UC1 }; };
template <typename T>
void eigen2cv(const Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> &src, Mat &dst) {
    dst.create(static_cast<int>(src.rows()), static_cast<int>(src.cols()), stub_type<T>::value);
    for (int y = 0; y < dst.rows; ++y) std::memcpy(dst.data + y * dst.step, src.data() + static_cast<size_t>(y) * src.cols(), sizeof(T) * src.cols());
}
}  // namespace cv
#endif
