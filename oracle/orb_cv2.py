"""TEST INFRASTRUCTURE / CPU BASELINE -- never imported by the product path.

`ORBextractor::operator()` composed from OpenCV primitives the way the reference composes them
(src/orbslam/ORBextractor.cc:752-847, 1019-1122: cv::resize, cv::copyMakeBorder, cv::FAST per 30-px cell with
the threshold retry, cv::GaussianBlur, IC angle, rBRIEF), with the oracle's quad tree.  It is the closest
thing to the reference's CPU path that runs in this image (cv2 4.13 wheel; the C++ OpenCV the reference links
is absent), so bench.py times it as the CPU baseline; tests/test_oracle_orb.py checks that it and the numpy
restatement agree bit for bit.
"""
from __future__ import annotations

import numpy as np

from oracle import orb_oracle as O


def extract(gray: np.ndarray, nfeatures: int = 2000):
    import cv2
    p = O.ExtractorParams(nfeatures)
    t = O.Tables(p)
    pattern = O.load_pattern()
    h, w = gray.shape
    levels = []
    for lvl, (lw, lh) in enumerate(O.level_sizes(w, h, t)):
        levels.append(gray if lvl == 0 else cv2.resize(levels[-1], (lw, lh), interpolation=cv2.INTER_LINEAR))
    det_ini = cv2.FastFeatureDetector_create(p.ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_min = cv2.FastFeatureDetector_create(p.min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps, descs = [], []
    for lvl, im in enumerate(levels):
        lh, lw = im.shape
        cells, (mbx, mby, Mbx, Mby) = O.level_cells(lw, lh)
        cx, cy, cr = [], [], []
        for (x0, y0, x1, y1) in cells:
            sub = np.ascontiguousarray(im[y0:y1, x0:x1])
            k = det_ini.detect(sub)
            if not k:
                k = det_min.detect(sub)
            for q in k:
                cx.append(q.pt[0] + x0 - mbx)
                cy.append(q.pt[1] + y0 - mby)
                cr.append(q.response)
        if not cx:
            continue
        sel = O.distribute_octtree(np.array(cx), np.array(cy), np.array(cr, np.float32), mbx, Mbx, mby, Mby,
                                   t.per_level[lvl])
        bordered = cv2.copyMakeBorder(im, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
        blur = cv2.GaussianBlur(im.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        sc = t.scale[lvl]
        size = np.float32(int(np.float32(31) * sc))
        for s in sel:
            x, y = int(cx[s]) + mbx, int(cy[s]) + mby
            ang = O.ic_angle(bordered, x, y, t.umax)
            descs.append(O.orb_descriptor(blur, x, y, ang, pattern))
            fx, fy = np.float32(x), np.float32(y)
            if lvl:
                fx, fy = np.float32(fx * sc), np.float32(fy * sc)
            kps.append([fx, fy, size, ang, cr[s], lvl, -1])
    return np.array(kps, np.float32).reshape(-1, 7), np.array(descs, np.uint8).reshape(-1, 32)
