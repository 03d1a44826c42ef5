"""Generates tests/golden/features.npz: golden vectors for the feature spaces (hog / ehog / lbp / whi /
histeq), made by
  - cv2 4.13.0 for the OpenCV-owned primitives (cv::Sobel as GradientFilter calls it, cv::equalizeHist,
    cv::dft as WhiteningFilter calls it), and
  - the reference's OWN compiled sources (oracle/_ref: GradientBinningFilter.cpp, HistogramFilter.cpp,
    SpatialHistogramFilter.cpp, HogFilter.cpp, ExtendedHogFilter.cpp, LbpFilter.cpp/.hpp) for the in-repo filters,
on seeded inputs.  Inputs that cannot be regenerated from seeds are stored next to the outputs."""
import os
import sys
import zlib

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from featuredetection_b200 import capi, synthetic as syn  # noqa: E402
from oracle import fdoracle as fo  # noqa: E402

assert fo.ref_available(), "oracle/_ref is not built (needs /root/reference)"
cv2.setNumThreads(1)
out = {}
frame = syn.synthetic_frame(0)
rng = np.random.default_rng(4242)
img = np.ascontiguousarray(frame[60:180, 200:360])                      # smooth + noise
noise = rng.integers(0, 256, (64, 80), dtype=np.uint8)
out["noise"] = noise

# --- cv2: Sobel as GradientFilter.cpp:54-55 calls it -------------------------------------------------------------
for name, src in (("img", img), ("noise", noise)):
    for k, sc in ((1, 0.5), (3, 0.125)):
        gx = cv2.Sobel(src, cv2.CV_8U, 1, 0, ksize=k, scale=sc, delta=127)
        gy = cv2.Sobel(src, cv2.CV_8U, 0, 1, ksize=k, scale=sc, delta=127)
        out["sobel_%s_k%d" % (name, k)] = np.stack([gx, gy], axis=-1)

# --- cv2: equalizeHist + the whitening chain on patches ----------------------------------------------------------
PATCH_SIZES = [(20, 20), (24, 24), (32, 16), (32, 24), (16, 24)]
pp = []
for i in range(60):
    w, h = PATCH_SIZES[i % 5]
    src = noise if i % 3 == 2 else img
    x = int(rng.integers(0, src.shape[1] - w)); y = int(rng.integers(0, src.shape[0] - h))
    pp.append((i % 3 == 2, x, y, w, h))
out["patch_boxes"] = np.array(pp, np.int32)
for i, (is_noise, x, y, w, h) in enumerate(pp):
    p = np.ascontiguousarray((noise if is_noise else img)[y:y + h, x:x + w])
    out["histeq_%d" % i] = cv2.equalizeHist(p)
    F = fo.whitening_filter(w, h)
    X = cv2.dft(p.astype(np.float32), flags=cv2.DFT_SCALE | cv2.DFT_COMPLEX_OUTPUT)
    X = X * F[:, :, None]
    yv = cv2.dft(X, flags=cv2.DFT_INVERSE | cv2.DFT_REAL_OUTPUT)
    u8 = np.clip(np.rint(yv + np.float32(127)), 0, 255).astype(np.uint8)          # convertTo(CV_8U, 1, 127)
    out["whi_u8_%d" % i] = u8
    eq = cv2.equalizeHist(u8)
    f = eq.astype(np.float32) * np.float32(1.0 / 127.5) + np.float32(-1.0)
    nrm = cv2.norm(f, cv2.NORM_L2)
    out["whi_vec_%d" % i] = (f * np.float32(1.0 / (nrm + np.float64(np.float32(1e-4))))).ravel()

# --- compiled reference: binning LUT checksums, LBP codes, patch histograms --------------------------------------
for bins, sg in ((9, 0), (18, 1), (8, 1)):
    one, two = fo.ref_gradient_bin_luts(bins, sg)
    out["lut_crc_%d_%d" % (bins, sg)] = np.array([zlib.crc32(one.tobytes()), zlib.crc32(two.tobytes())], np.uint32)
    out["lut_sample_%d_%d" % (bins, sg)] = two[::251].copy()
for t in range(4):
    out["lbp_%d" % t] = fo.ref_lbp(img, t)

CASES = [
    dict(kind="hog"), dict(kind="hog", interpolate_cells=True), dict(kind="hog", interpolate_bins=False, gradient_kernel=3),
    dict(kind="hog", cell_size=6, normalization="l2hys"), dict(kind="hog", normalization="l1sqrt", interpolate_cells=True),
    dict(kind="hog", block_size=2), dict(kind="hog", block_size=2, signed_and_unsigned=True, bins=8, signed_gradients=True),
    dict(kind="ehog"), dict(kind="ehog", signed_and_unsigned=True, bins=18, signed_gradients=True, interpolate_cells=True),
    dict(kind="lbp", cell_size=10), dict(kind="lbp", lbp_type="lbp8", cell_size=10, interpolate_cells=True),
    dict(kind="lbp", lbp_type="lbp4", block_size=2, concatenate=True, normalization="l2hys"),
    dict(kind="lbp", lbp_type="lbp4rotated", block_size=2, concatenate=False, normalization="l1norm"),
]
out["hist_cases"] = np.array([repr(c) for c in CASES])
boxes = []
for ci, kw in enumerate(CASES):
    d = syn.feature_desc(**kw)
    for (pw, ph) in ((20, 20), (30, 30), (32, 16)):
        F = fo.Features(d, pw, ph)
        fl = F.filter_layer(img)   # layer filters: pinned separately (Sobel vs cv2, LUT/LBP vs compiled reference)
        bins = fo.lib().fdo_lbp_bins(d.lbp_type) if d.kind == capi.FDB_FEATURE_LBP else d.bins
        vecs = []
        for i in range(6):
            x = int(rng.integers(0, img.shape[1] - pw)); y = int(rng.integers(0, img.shape[0] - ph))
            boxes.append((ci, pw, ph, x, y))
            vecs.append(fo.ref_patch_histogram(d, bins, fl, x, y, pw, ph))
        out["hist_%d_%dx%d" % (ci, pw, ph)] = np.stack(vecs)
out["hist_boxes"] = np.array(boxes, np.int32)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "features.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")
