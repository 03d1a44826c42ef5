"""Generates tests/golden/sdm.npz: golden vectors for the supervised-descent path (BASELINE configs[4]), made by
  - the reference's OWN vendored VLFeat HOG (libSupervisedDescent/src/superviseddescent/hog.c compiled unmodified into
    oracle/_ref) on seeded float32 30x30 patches,
  - cv2 4.13.0 for the OpenCV-owned primitives: cv::resize on CV_32F (DescriptorExtractor.hpp:182) and cv::gemm on CV_32F
    (the MatExpr `features * R.rowRange(...) + R.row(...)` of SdmLandmarkModel.hpp:241),
  - the oracle's whole fit with the reference's in-repo model detect-landmarks/share/models/
    SDM_Model_HOG_Zhenhua_22072014.txt (real weights; the model file itself stays in /root/reference) and with the seeded
    synthetic 68-landmark model.
Run here (needs /root/reference and cv2); the .npz travels to the GPU box."""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from featuredetection_b200 import synthetic as syn  # noqa: E402
from oracle import fdoracle as fo  # noqa: E402

assert fo.ref_available(), "oracle/_ref is not built (needs /root/reference)"
cv2.setNumThreads(1)
# The reference pins OpenCV 2.4.3, whose float resize is SSE2 mul + add. cv2 4.13's dispatched AVX2 build contracts those
# into FMA3 (results differ in the last bit for ~60 % of the pixels); its baseline SSE3 code path - selected by
# setUseOptimized(False) - has the 2.4.3 arithmetic and is what the golden vectors record.
cv2.setUseOptimized(False)
MODEL = "/root/reference/detect-landmarks/share/models/SDM_Model_HOG_Zhenhua_22072014.txt"
out = {}
rng = np.random.default_rng(777)

# --- hog.c: 24 patches (smooth crops, noise, flat, a step edge) ---------------------------------------------------
frame = syn.synthetic_frame(0)
patches = []
for i in range(24):
    if i < 12:
        y, x = rng.integers(0, 440), rng.integers(0, 600)
        p = frame[y:y + 30, x:x + 30].astype(np.float32)
    elif i < 20:
        p = rng.integers(0, 256, (30, 30)).astype(np.float32)
    elif i == 20:
        p = np.full((30, 30), 97, np.float32)
    elif i == 21:
        p = np.zeros((30, 30), np.float32); p[:, 15:] = 255
    else:
        p = (rng.random((30, 30)) * 255).astype(np.float32)      # non-integer pixel values (after a resize)
    patches.append(p)
patches = np.stack(patches)
out["hog_patches"] = patches
out["hog_ref"] = np.stack([fo.vlhog_uoctti(p, 10, 9, use_ref=True) for p in patches])

# --- cv2.resize CV_32F INTER_LINEAR to 30x30 from the window sizes the cascade produces -----------------------------
sizes = [12, 18, 24, 30, 36, 42, 48, 54, 60, 66, 72, 90, 120]
out["resize_sizes"] = np.array(sizes)
for s in sizes:
    src = rng.integers(0, 256, (s, s)).astype(np.float32)
    out["resize_src_%d" % s] = src
    out["resize_dst_%d" % s] = cv2.resize(src, (30, 30), interpolation=cv2.INTER_LINEAR)

# --- cv2.gemm CV_32F: 1 x K times K x N plus bias -------------------------------------------------------------------
K, N = 15 * 279, 30
f = (rng.random((1, K)) * 0.4).astype(np.float32)
R = (rng.standard_normal((K + 1, N)) * 1e-3).astype(np.float32)
out["gemm_f"], out["gemm_R"] = f, R
out["gemm_out"] = cv2.gemm(f, R[:K], 1.0, R[K:K + 1], 1.0)

# --- whole fits ----------------------------------------------------------------------------------------------------
frames = syn.synthetic_frames(0, 4)
boxes = np.array([[220, 140, 200, 200], [100, 60, 260, 260], [300, 200, 150, 150], [5, 5, 200, 200], [400, 250, 230, 220]], np.int32)
out["fit_boxes"] = boxes
real = fo.Sdm(path=MODEL)
synth = fo.Sdm(syn.make_sdm(68, 5, 500))
for name, m in (("real", real), ("synth", synth)):
    shapes0, shapes, feat_sums = [], [], []
    for k in range(4):
        for b in boxes:
            s0 = m.align_rigid(b)
            try:
                s1, feats = m.optimize(frames[k], s0, want_features=True)
                fs = feats.astype(np.float64).sum(axis=1)
            except RuntimeError:
                s1, fs = np.full_like(s0, np.nan), np.full(m.steps, np.nan)
            shapes0.append(s0); shapes.append(s1); feat_sums.append(fs)
    out["fit_%s_start" % name] = np.stack(shapes0)
    out["fit_%s_shapes" % name] = np.stack(shapes)
    out["fit_%s_feature_sums" % name] = np.stack(feat_sums)
    print(name, "faces", len(shapes), "failed", int(np.isnan(np.stack(shapes)[:, 0]).sum()))
# the real model's mean and a checksum of its regressors pin the text loader
rm = real.to_model()
out["real_mean"] = rm.mean
out["real_reg_sums"] = np.array([r.astype(np.float64).sum() for r in rm.regressors])
out["real_reg_corner"] = np.stack([r[-2:, :4] for r in rm.regressors])

np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sdm.npz"), **out)
print("wrote sdm.npz")
