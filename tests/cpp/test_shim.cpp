// Runs the reference's own BayesianSegNet tests (tests/test_bayesian_segnet.cpp:138-168: InitializationTest,
// SegmentationTest) and the two operators through the C++ shim classes of integration/ -- the unmodified shim sources,
// compiled against integration/stubs (OpenCV / Eigen are absent from the image) and linked to libsivo_b200.so.
//
//   test_shim init  <prototxt> <caffemodel>                         ctor error behaviour only (no GPU needed)
//   test_shim run   <prototxt> <caffemodel> <bgr.bin> <gray.bin> <out_dir>
//
// *.bin = int32 rows, int32 cols, int32 channels, then the pixels.  `run` dumps classes / confidence / entropy, the blended
// segmentation image, keypoints, descriptors and the pyramid levels for tests/test_shim.py to compare with the Python mirror
// of the same C-ABI (bit-identical) and with cv2 (LUT + addWeighted).
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "bayesian_segnet/bayesian_segnet.hpp"
#include "orbslam/ORBextractor.h"

using namespace SIVO;

static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

template <class F> static bool throws_invalid_argument(F &&f) {
    try { f(); } catch (const std::invalid_argument &) { return true; } catch (...) { return false; }
    return false;
}

static cv::Mat read_bin(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    int hdr[3] = {0, 0, 0};
    f.read(reinterpret_cast<char *>(hdr), sizeof hdr);
    cv::Mat m(hdr[0], hdr[1], hdr[2] == 3 ? CV_8UC3 : CV_8UC1);
    f.read(reinterpret_cast<char *>(m.data), static_cast<std::streamsize>(m.total() * m.elemSize()));
    if (!f) throw std::runtime_error("cannot read " + path);
    return m;
}
static void dump(const std::string &path, const void *p, size_t bytes) {
    std::ofstream f(path, std::ios::binary);
    f.write(static_cast<const char *>(p), static_cast<std::streamsize>(bytes));
}

// tests/test_bayesian_segnet.cpp:138-150
static void InitializationTest(const std::string &model, const std::string &weights, bool construct_valid) {
    std::string bad_model, bad_weights;
    EXPECT(throws_invalid_argument([&] { BayesianSegNet m(BayesianSegNetParams(model, bad_weights)); }));
    EXPECT(throws_invalid_argument([&] { BayesianSegNet m(BayesianSegNetParams(bad_model, weights)); }));
    EXPECT(throws_invalid_argument([&] { BayesianSegNet m(BayesianSegNetParams(bad_model, bad_weights)); }));
    if (construct_valid) {
        bool ok = true;
        try { BayesianSegNet m(BayesianSegNetParams(model, weights)); } catch (...) { ok = false; }
        EXPECT(ok);
    }
}

int main(int argc, char **argv) {
    if (argc < 4) { std::printf("usage: test_shim init|run <prototxt> <caffemodel> [bgr.bin gray.bin out_dir]\n"); return 2; }
    const std::string mode = argv[1], model = argv[2], weights = argv[3];
    if (mode == "init") {
        InitializationTest(model, weights, false);
        std::printf(failures ? "SHIM INIT FAILED\n" : "SHIM INIT OK\n");
        return failures ? 1 : 0;
    }
    if (argc < 7) return 2;
    const std::string out = argv[6];
    InitializationTest(model, weights, true);
    {   // tests/test_bayesian_segnet.cpp:152-168 (SegmentationTest) + generateSegmentedImage (bayesian_segnet.cpp:362-389)
        cv::Mat image = read_bin(argv[4]);
        BayesianSegNet net{BayesianSegNetParams{model, weights}};
        cv::Size g = net.getInputGeometry();
        MatXu classes;
        MatXd confidence, entropy;
        net.segmentImage(image, classes, confidence, entropy);
        EXPECT(classes.size() == g.height * g.width);
        EXPECT(confidence.size() == g.height * g.width);
        EXPECT(entropy.size() == g.height * g.width);
        cv::Mat seg = net.generateSegmentedImage(classes, image);
        EXPECT(seg.rows == g.height && seg.cols == g.width && seg.type() == CV_8UC3);
        cv::Mat ent_img = net.generateEntropyImage(entropy);
        EXPECT(ent_img.rows == g.height && ent_img.cols == g.width);
        dump(out + "/classes.bin", classes.data(), classes.size());
        dump(out + "/confidence.bin", confidence.data(), confidence.size() * sizeof(double));
        dump(out + "/entropy.bin", entropy.data(), entropy.size() * sizeof(double));
        cv::Mat segc = seg.clone();
        dump(out + "/segmented.bin", segc.data, segc.total() * 3);
        cv::Mat entc = ent_img.clone();
        dump(out + "/entropy_image.bin", entc.data, entc.total() * sizeof(double));
        std::printf("geometry %d %d\n", g.width, g.height);
    }
    {   // ORBextractor::operator() (ORBextractor.cc:1019-1083) with the public pyramid Frame::ComputeStereoMatches reads
        cv::Mat gray = read_bin(argv[5]);
        ORBextractor ex(1000, 1.2f, 8, 20, 7);
        std::vector<cv::KeyPoint> kps;
        cv::Mat desc;
        ex(gray, cv::Mat(), kps, desc);
        EXPECT(!kps.empty() && desc.rows == static_cast<int>(kps.size()) && desc.cols == 32);
        EXPECT(ex.GetLevels() == 8 && ex.mvImagePyramid.size() == 8);
        EXPECT(ex.mvImagePyramid[0].rows == gray.rows && ex.mvImagePyramid[0].cols == gray.cols);
        // the per-level tables Frame / ORBmatcher read through the getters (ORBextractor.cc:420-438)
        std::vector<float> sf = ex.GetScaleFactors(), isf = ex.GetInverseScaleFactors(), s2 = ex.GetScaleSigmaSquares(),
                           is2 = ex.GetInverseScaleSigmaSquares();
        EXPECT(sf.size() == 8 && isf.size() == 8 && s2.size() == 8 && is2.size() == 8);
        EXPECT(std::fabs(ex.GetScaleFactor() - 1.2) < 1e-6 && sf[0] == 1.0f && s2[0] == 1.0f);
        for (size_t l = 1; l < sf.size() && l < 8; ++l) {
            EXPECT(std::fabs(sf[l] - sf[l - 1] * 1.2f) <= 2e-6f * sf[l] && std::fabs(s2[l] - sf[l] * sf[l]) <= 2e-6f * s2[l]);
            EXPECT(std::fabs(isf[l] * sf[l] - 1.f) < 2e-6f && std::fabs(is2[l] * s2[l] - 1.f) < 2e-6f);
        }
        dump(out + "/keypoints.bin", kps.data(), kps.size() * sizeof(cv::KeyPoint));
        cv::Mat d = desc.clone();
        dump(out + "/descriptors.bin", d.data, d.total());
        for (int l = 0; l < 8; ++l) {
            cv::Mat lv = ex.mvImagePyramid[l].clone();
            dump(out + "/level" + std::to_string(l) + ".bin", lv.data, lv.total());
            std::printf("level %d %d %d\n", l, lv.cols, lv.rows);
        }
        std::vector<cv::KeyPoint> none;
        cv::Mat nodesc;
        ex(cv::Mat(), cv::Mat(), none, nodesc);  // empty image: returns silently (:1023-1024)
        EXPECT(none.empty());
    }
    std::printf(failures ? "SHIM RUN FAILED\n" : "SHIM RUN OK\n");
    return failures ? 1 : 0;
}
