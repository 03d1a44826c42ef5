"""Generates tests/golden/cv2_primitives.npz with OpenCV (cv2 4.13.0 in this image): the pinned
reference for the OpenCV-owned primitives of the path (cv::resize INTER_LINEAR at
ImagePyramid.cpp:177 and cv::pyrDown at :186). Run once here; the .npz is committed so the check
does not depend on cv2 being importable (or bit-stable across CPUs) where the tests run."""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from featuredetection_b200 import synthetic as syn  # noqa: E402

cv2.setNumThreads(1)
rng = np.random.default_rng(2024)
out = {"cv2_version": np.array(cv2.__version__)}
cases = []
for i, (sw, sh, dw, dh) in enumerate([(64, 48, 59, 44), (97, 61, 89, 56), (128, 96, 64, 48), (33, 35, 17, 31),
                                      (200, 150, 109, 82), (50, 40, 50, 40), (31, 7, 13, 3), (120, 90, 110, 83)]):
    img = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    out["resize_src_%d" % i] = img
    out["resize_dst_%d" % i] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    cases.append((sw, sh, dw, dh))
out["resize_cases"] = np.array(cases)
for i, (w, h) in enumerate([(64, 48), (97, 61), (33, 35), (5, 4), (2, 9), (1, 6), (200, 3), (75, 111)]):
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    out["pyrdown_src_%d" % i] = img
    out["pyrdown_dst_%d" % i] = cv2.pyrDown(img)
out["pyrdown_n"] = np.array(8)
# one real pyramid chain on a synthetic frame crop: FaceFrontal octave offsets 1 and 6, two halvings
f = syn.synthetic_frame(0)[:240, :320]
out["chain_src"] = f
for i in (1, 6):
    q = 0.5 ** (i / 8.0)
    w, h = int(np.rint(320 * q)), int(np.rint(240 * q))
    a = cv2.resize(f, (w, h), interpolation=cv2.INTER_LINEAR)
    b = cv2.pyrDown(a)
    c = cv2.pyrDown(b)
    out["chain_%d_0" % i], out["chain_%d_1" % i], out["chain_%d_2" % i] = a, b, c
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv2_primitives.npz"), **out)
print("written", len(out), "arrays")
