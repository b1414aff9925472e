"""Runs the extractor alone a few times (for `ncu --metrics gpu__time_duration.sum`: the per-kernel times of one ORB call)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sivo_b200 import ORBextractor  # noqa: E402
from sivo_b200.synth import bgr_to_gray, stereo_frame  # noqa: E402

left, _ = stereo_frame(0)
gray = np.ascontiguousarray(bgr_to_gray(left)[11:11 + 352, 109:109 + 1024])
ext = ORBextractor(2000, 1.2, 8, 20, 7)
for _ in range(4):
    kps, desc = ext(gray, None)
print(len(kps), ext.last_timing())
