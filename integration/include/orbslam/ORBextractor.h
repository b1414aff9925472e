/* Drop-in replacement for SIVO's include/orbslam/ORBextractor.h (:46-125): identical public API, including the
 * public std::vector<cv::Mat> mvImagePyramid whose level-k Mat is a (19,19)-offset ROI of a bordered host buffer
 * (read by Frame::ComputeStereoMatches, Frame.cc:451,546-548,562,568-570). */
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <opencv2/core/core.hpp>
#include <vector>

#include "sivo_b200.h"

namespace SIVO {

class ORBextractor {
 public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints,
                    cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    double inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    std::vector<cv::Mat> mvImagePyramid;

 protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    sivo_orb_t *handle = nullptr;
    std::vector<cv::Mat> bordered;  // owners of the level buffers mvImagePyramid points into
};
}  // namespace SIVO
#endif
