"""Generates tests/golden/ref_classifiers.npz by RUNNING THE REFERENCE'S OWN SOURCES (oracle/_ref:
unmodified HistEq64Filter.cpp, IImg.cpp, WvmClassifier.cpp, SvmClassifier.cpp, RbfKernel.hpp,
Probabilistic*Classifier.cpp, OverlapElimination.cpp compiled from /root/reference) on seeded inputs.
These are the golden vectors that pin the oracle and the CUDA path where /root/reference is absent
(the GPU box). Inputs are regenerated from seeds by featuredetection_b200.synthetic; only small
arrays are stored."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from featuredetection_b200 import capi, synthetic as syn  # noqa: E402
from oracle import fdoracle as fo  # noqa: E402

assert fo.ref_available(), "oracle/_ref is not built (needs /root/reference)"
out = {}
det_kw, wvm, svm = syn.landmark_models("FaceFrontal")
_, wvm_ne, _ = syn.landmark_models("FaceFrontal", profile="no-exit")
frame = syn.synthetic_frame(0)
_, layers = fo.pyramid(frame, det_kw["incremental_scale_factor"], det_kw["min_scale_factor"], det_kw["max_scale_factor"])
rng = np.random.default_rng(77)
raw, eq = [], []
for k in range(96):
    idx, scale, img = layers[k % len(layers)]
    y = int(rng.integers(0, img.shape[0] - 20)); x = int(rng.integers(0, img.shape[1] - 20))
    p = np.ascontiguousarray(img[y:y + 20, x:x + 20])
    raw.append(p)
    eq.append(fo.hq64(img[y:y + 20, x:x + 20], use_ref=True))
raw, eq = np.stack(raw), np.stack(eq)
eq_extra = np.stack([np.zeros((20, 20), np.uint8), np.full((20, 20), 255, np.uint8),
                     rng.integers(0, 256, (20, 20), dtype=np.uint8), np.tile(np.arange(20, dtype=np.uint8) * 13, (20, 1))])
out["hq64_in"], out["hq64_out"] = raw, eq
patches = np.concatenate([eq, eq_extra]).reshape(-1, 400)
out["patches"] = patches
for tag, model in (("realistic", wvm), ("noexit", wvm_ne)):
    lv, fo_, pr, pos = fo.Wvm(model, use_ref=True).eval(patches)
    out["wvm_%s_level" % tag], out["wvm_%s_fout" % tag] = lv, fo_
    out["wvm_%s_prob" % tag], out["wvm_%s_pos" % tag] = pr, pos
d, p, q = fo.Svm(svm, use_ref=True).eval(patches)
out["svm_dist"], out["svm_prob"], out["svm_pos"] = d, p, q
# whole frames through the reference's own classes (ref_driver.cpp: ref_detect_frame)
wr, sr = fo.Wvm(wvm, use_ref=True), fo.Svm(svm, use_ref=True)
for k in (0, 1):
    for stage, name in ((capi.FDB_STAGE_WVM, "wvm"), (capi.FDB_STAGE_OE, "oe"), (capi.FDB_STAGE_SVM, "svm"), (capi.FDB_STAGE_NMS, "nms")):
        r = fo.ref_detect_frame(det_kw, wr, sr, syn.synthetic_frame(k), stage=stage, want_dense=(stage == capi.FDB_STAGE_WVM))
        out["frame%d_%s_windows" % (k, name)] = r["det_windows"]
        if r["dense"] is not None:
            out["frame%d_dense_level" % k] = r["dense"]["level"].astype(np.int16)
            out["frame%d_dense_fout" % k] = r["dense"]["fout"]
# overlap elimination on random candidate sets (distinct probabilities)
for i in range(4):
    n = int(rng.integers(5, 120))
    cx = rng.integers(0, 200, n).astype(np.int32); cy = rng.integers(0, 150, n).astype(np.int32)
    w = rng.choice([135, 147, 160, 174], n).astype(np.int32); p = rng.permutation(n).astype(np.float64) / n
    dist, ratio = [(5.0, 0.0), (0.5, 0.7), (12.0, 0.9), (30.0, 0.0)][i]
    keep = np.zeros(n, np.int32)
    m = fo.ref().ref_overlap_eliminate(dist, ratio, n, cx.ctypes.data, cy.ctypes.data, w.ctypes.data, p.ctypes.data, keep.ctypes.data)
    out["oe%d_in" % i] = np.stack([cx, cy, w]).astype(np.int32)
    out["oe%d_prob" % i] = p
    out["oe%d_param" % i] = np.array([dist, ratio])
    out["oe%d_keep" % i] = keep[:m].copy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_classifiers.npz"), **out)
print("written", len(out), "arrays")
