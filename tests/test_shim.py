"""The C++ drop-in classes of integration/ (SIVO::BayesianSegNet, SIVO::ORBextractor with the reference's signatures),
compiled from their unmodified sources and run: the reference's own InitializationTest / SegmentationTest
(tests/test_bayesian_segnet.cpp:138-168), `generateSegmentedImage` (bayesian_segnet.cpp:362-389, row a11) against cv2, and
both operators bit-identical to the Python mirror of the same C-ABI."""
import os
import subprocess

import numpy as np
import pytest

import shim_build
from conftest import make_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_bin(path, img):
    img = np.ascontiguousarray(img)
    hdr = np.array([img.shape[0], img.shape[1], 1 if img.ndim == 2 else img.shape[2]], np.int32)
    with open(path, "wb") as f:
        f.write(hdr.tobytes())
        f.write(img.tobytes())


def test_shim_compiles_and_constructor_throws_like_the_reference(tmp_path):
    """g++ -Wall -Werror over the shim sources; empty model / weights paths -> std::invalid_argument
    (bayesian_segnet.cpp:80-89, pinned by tests/test_bayesian_segnet.cpp:146-149).  Needs no GPU."""
    exe = shim_build.build()
    r = subprocess.run([exe, "init", os.path.join(ROOT, "configs", "bayesian_segnet_basic.prototxt"), str(tmp_path / "w.caffemodel")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "SHIM INIT OK" in r.stdout, r.stdout + r.stderr


# class colours of BayesianSegNet::BayesianSegNet (bayesian_segnet.cpp:91-117), BGR; classes >= 14 stay black
COLOURS = [(128, 64, 128), (232, 35, 244), (69, 69, 69), (156, 102, 102), (153, 153, 153), (30, 170, 250), (0, 220, 220),
           (35, 142, 107), (152, 251, 152), (180, 130, 70), (60, 20, 220), (142, 0, 0), (70, 0, 0), (32, 11, 119)]


@pytest.mark.gpu
def test_shim_runs_the_reference_tests_and_equals_the_python_mirror(tmp_path, kitti_bgr):
    import cv2
    from sivo_b200 import BayesianSegNet, BayesianSegNetParams, ORBextractor
    from sivo_b200.synth import bgr_to_gray
    exe = shim_build.build()
    H, W, T = 96, 256, 3
    net, w, proto, model = make_model(tmp_path, "basic", T=T, H=H, W=W)
    img = np.ascontiguousarray(kitti_bgr[60:60 + H + 20, 200:200 + W + 30])   # larger than the net input: centre-cropped
    gray = np.ascontiguousarray(bgr_to_gray(kitti_bgr)[11:11 + 352, 109:109 + 1024])
    write_bin(tmp_path / "bgr.bin", img)
    write_bin(tmp_path / "gray.bin", gray)
    out = tmp_path / "out"
    out.mkdir()
    env = dict(os.environ, SIVO_B200_SEED="1234")
    r = subprocess.run([exe, "run", proto, model, str(tmp_path / "bgr.bin"), str(tmp_path / "gray.bin"), str(out)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "SHIM RUN OK" in r.stdout, r.stdout + r.stderr
    seg = BayesianSegNet(BayesianSegNetParams(proto, model), seed=1234)
    seg.set_frame(0)
    cls, conf, ent = seg.segmentImage(img)
    assert np.array_equal(np.fromfile(out / "classes.bin", np.uint8).reshape(H, W), cls)
    assert np.array_equal(np.fromfile(out / "confidence.bin", np.float64).reshape(H, W), conf)
    assert np.array_equal(np.fromfile(out / "entropy.bin", np.float64).reshape(H, W), ent)
    # generateSegmentedImage: LUT colourise + 50/50 blend with the centre-cropped input, against real cv2
    lut = np.zeros((256, 1, 3), np.uint8)
    lut[:14, 0] = COLOURS
    y0, x0 = img.shape[0] // 2 - H // 2, img.shape[1] // 2 - W // 2
    crop = np.ascontiguousarray(img[y0:y0 + H, x0:x0 + W])
    want = cv2.addWeighted(cv2.LUT(cv2.cvtColor(cls, cv2.COLOR_GRAY2BGR), lut), 0.5, crop, 0.5, 0)
    got = np.fromfile(out / "segmented.bin", np.uint8).reshape(H, W, 3)
    assert np.array_equal(got, want)
    ent_img = np.fromfile(out / "entropy_image.bin", np.float64).reshape(H, W)
    assert np.allclose(ent_img, cv2.normalize(ent, None, 0.0, 1.0, cv2.NORM_MINMAX, cv2.CV_64FC1), atol=1e-12)
    # ORBextractor::operator(): keypoints (cv::KeyPoint records), descriptors and the public pyramid
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    kps, desc = ex(gray, None, want_pyramid=True)
    pyr = ex._bordered  # the bordered level buffers mvImagePyramid points into
    got_kp = np.fromfile(out / "keypoints.bin", kps.dtype)
    assert len(got_kp) == len(kps) and got_kp.tobytes() == kps.tobytes()
    assert np.array_equal(np.fromfile(out / "descriptors.bin", np.uint8).reshape(-1, 32), desc)
    for l in range(8):
        h, w_ = pyr[l].shape[0] - 38, pyr[l].shape[1] - 38
        lv = np.fromfile(out / f"level{l}.bin", np.uint8).reshape(h, w_)
        assert np.array_equal(lv, pyr[l][19:19 + h, 19:19 + w_]), l
