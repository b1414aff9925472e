/* Shim translation unit: SIVO::ORBextractor over the C-ABI.  Replaces src/orbslam/ORBextractor.cc in the `orbslam`
 * source list (CMakeLists.txt:96). */
#include "orbslam/ORBextractor.h"

#include <cassert>
#include <cstdlib>
#include <stdexcept>

namespace SIVO {

static_assert(sizeof(cv::KeyPoint) == sizeof(sivo_keypoint), "sivo_keypoint must stay bit-compatible with cv::KeyPoint");

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels) {
    const char *dev = std::getenv("SIVO_B200_DEVICE");
    if (sivo_orb_create(_nfeatures, _scaleFactor, _nlevels, _iniThFAST, _minThFAST, dev ? std::atoi(dev) : 0, &handle) != SIVO_OK)
        throw std::runtime_error(sivo_last_error());
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    sivo_orb_tables(handle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), nullptr);
    mvImagePyramid.resize(nlevels);
    bordered.resize(nlevels);
}

ORBextractor::~ORBextractor() { sivo_orb_destroy(handle); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray, std::vector<cv::KeyPoint> &_keypoints,
                              cv::OutputArray _descriptors) {
    if (_image.empty()) return;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);
    std::vector<uint8_t *> ptrs(nlevels);
    std::vector<size_t> strides(nlevels);
    for (int l = 0; l < nlevels; ++l) {
        int w = 0, h = 0;
        sivo_orb_level_size(handle, image.rows, image.cols, l, &w, &h);
        bordered[l].create(h + 38, w + 38, CV_8UC1);
        mvImagePyramid[l] = bordered[l](cv::Rect(19, 19, w, h));
        ptrs[l] = bordered[l].data;
        strides[l] = bordered[l].step;
    }
    const int cap = nfeatures + 4 * nlevels + 64;
    _keypoints.resize(cap);
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0;
    int rc = sivo_orb_run(handle, image.data, image.rows, image.cols, image.step,
                          reinterpret_cast<sivo_keypoint *>(_keypoints.data()), cap, &n, desc.data, ptrs.data(), strides.data());
    if (rc != SIVO_OK) throw std::runtime_error(sivo_last_error());
    _keypoints.resize(n);
    if (n == 0) _descriptors.release();
    else desc.rowRange(0, n).copyTo(_descriptors);
}
}  // namespace SIVO
