"""The C++ adapters (adapters/fdb200_adapters.hpp) EXECUTED on the GPU behind the reference's own, unchanged interface headers:
oracle/_ref/run_adapters (built by oracle/Makefile from adapters/run_adapters.cpp where /root/reference is mounted, against the
cv::Mat stand-in of oracle/shim) constructs B200SlidingWindowDetector / B200SingleDetector / the classifier adapters from MATLAB
model files and calls Detector::detect(Mat), detect(Mat, Rect), PyramidFeatureExtractor::extract(...) in all its forms and
ProbabilisticClassifier::getProbability; everything it prints is compared with the oracle."""
import math
import os
import subprocess

import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn
from featuredetection_b200.detector import load_wvm_mat, load_svm_mat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "oracle", "_ref", "run_adapters")


def _fnv(patch):
    h = 1469598103934665603
    for b in np.asarray(patch, np.uint8).ravel().tolist():
        h ^= b
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _fold(values):
    h = 0
    for v in values:
        h = (h * 31 + v) & 0xFFFFFFFFFFFFFFFF
    return h


def _geom(x, y, w, h):
    return (x * 7919 + y * 104729 + w * 13 + h) & 0xFFFFFFFFFFFFFFFF


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(RUNNER), reason="oracle/_ref/run_adapters not built (needs /root/reference at build time)")
def test_adapters_behind_the_reference_interfaces(built, tmp_path):
    from oracle import fdoracle as fo
    sio = pytest.importorskip("scipy.io")
    kw, wvm0, svm0 = syn.landmark_models("FaceFrontal")
    cpath, tpath, spath = (str(tmp_path / n) for n in ("wvm.mat", "thr.mat", "svm.mat"))
    syn.write_wvm_mat(wvm0, cpath, tpath, True)
    syn.write_svm_mat(svm0, 20, 20, spath)
    wvm, svm = load_wvm_mat(cpath, tpath), load_svm_mat(spath)   # the models exactly as the library reads them
    svm.logistic_a, svm.logistic_b = 0.00556, -2.95
    frame = syn.synthetic_frame(41)
    fpath = str(tmp_path / "frame.raw")
    frame.tofile(fpath)
    inc, mn, mx = kw["incremental_scale_factor"], kw["min_scale_factor"], kw["max_scale_factor"]
    r = subprocess.run([RUNNER, cpath, tpath, spath, fpath, "640", "480", repr(inc), repr(mn), repr(mx), "20", "20"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = {}
    for line in r.stdout.splitlines():
        t = line.split()
        out.setdefault(t[0], []).append(t[1:])
    wo, so = fo.Wvm(wvm), fo.Svm(svm)
    pw = ph = 20
    _, layers = fo.pyramid(frame, inc, mn, mx)
    by_index = {i: (s, img) for i, s, img in layers}
    L = fo.lib()

    def patch_of(layer_index, x, y):
        return fo.hq64(by_index[layer_index][1][y:y + ph, x:x + pw])

    def check_dets(tag, ref):
        items = out.get(tag + "_ITEM", [])
        assert int(out[tag][0][0]) == len(items) == len(ref), tag
        for it, d in zip(items, ref):
            assert [int(v) for v in it[:4]] == [int(d["center_x"]), int(d["center_y"]), int(d["width"]), int(d["height"])], tag
            assert abs(float(it[4]) - float(d["probability"])) <= 1e-9 and int(it[5]) == int(d["positive"]), tag
            assert (int(it[6]), int(it[7])) == (ph, pw) and int(it[8]) == _fnv(patch_of(int(d["layer"]), int(d["x"]), int(d["y"]))), tag

    check_dets("FIVE", fo.detect_frame(kw, wo, so, frame, stage=capi.FDB_STAGE_NMS)["detections"])
    check_dets("FIVE_ROI", fo.detect_frame(kw, wo, so, frame, stage=capi.FDB_STAGE_NMS, roi=(150, 100, 330, 300))["detections"])
    check_dets("WVM", fo.detect_frame(kw, wo, None, frame, stage=capi.FDB_STAGE_WVM)["detections"])

    def enumerate_patches(step_x, step_y, roi, first, last, step_layer):
        """DirectPyramidFeatureExtractor::extract (DirectPyramidFeatureExtractor.cpp:75-123) restated"""
        x0, y0, w0, h0 = roi
        if roi == (0, 0, 0, 0):
            w0, h0 = 640, 480
        else:
            x0, y0 = max(0, x0), max(0, y0)
            w0, h0 = min(640, roi[2] + x0) - x0, min(480, roi[3] + y0) - y0
        if first < 0:
            first = layers[0][0]
        if last < 0:
            last = layers[-1][0]
        sums, geo = [], []
        for k in range(0, len(layers), step_layer):
            index, scale, img = layers[k]
            if index < first:
                continue
            if index > last:
                break
            ow, oh = L.fdo_cvround(pw / scale), L.fdo_cvround(ph / scale)
            bx, by = L.fdo_cvround(x0 * scale), L.fdo_cvround(y0 * scale)
            ex, ey = L.fdo_cvround((x0 + w0) * scale), L.fdo_cvround((y0 + h0) * scale)
            y = by
            while y + ph < ey:
                x = bx
                while x + pw < ex:
                    sums.append(_fnv(fo.hq64(img[y:y + ph, x:x + pw])))
                    geo.append(_geom(L.fdo_cvround(x / scale) + ow // 2, L.fdo_cvround(y / scale) + oh // 2, ow, oh))
                    x += step_x
                y += step_y
        return len(sums), _fold(sums), _fold(geo)

    assert [int(v) for v in out["LAYERS"][0]] == [len(layers)]
    assert tuple(int(v) for v in out["EXTRACT_STEP"][0]) == enumerate_patches(2, 3, (100, 80, 300, 250), -1, -1, 2)
    assert tuple(int(v) for v in out["EXTRACT_LAYERS"][0]) == enumerate_patches(4, 4, (0, 0, 0, 0), layers[1][0], layers[3][0], 1)
    n_all, h_all, g_all = (int(v) for v in out["EXTRACT_ALL"][0])
    ref = fo.detect_frame(kw, wo, None, frame, stage=capi.FDB_STAGE_WVM, want_patches=True)
    assert n_all == ref["windows"] == 16185
    assert h_all == _fold(_fnv(p) for p in ref["patches"])
    assert (n_all, h_all, g_all) == enumerate_patches(1, 1, (0, 0, 0, 0), -1, -1, 1)

    # single windows of layer 2: inside the scan, last column + last row (beyond the scan's strict '<' bound), out of bounds
    index, scale, img = layers[2]
    lw, lh = img.shape[1], img.shape[0]
    xs = [pw // 2 + 3, lw - pw + pw // 2, lw - pw + pw // 2 + 1, pw // 2 - 1]
    ys = [ph // 2 + 5, lh - ph + ph // 2, ph // 2 + 5, ph // 2]
    ow, oh = L.fdo_cvround(pw / scale), L.fdo_cvround(ph / scale)
    for k, item in enumerate(out["SINGLE"]):
        x, y = xs[k] - pw // 2, ys[k] - ph // 2                    # DirectPyramidFeatureExtractor.cpp:125-131
        inside = x >= 0 and y >= 0 and x + pw <= lw and y + ph <= lh   # :135-136
        assert inside == (k < 2)
        if not inside:
            assert item[1] == "none"
            continue
        want = [L.fdo_cvround(x / scale) + ow // 2, L.fdo_cvround(y / scale) + oh // 2, ow, oh, _fnv(fo.hq64(img[y:y + ph, x:x + pw]))]
        assert [int(v) for v in item[1:]] == want

    # extract(x, y, w, h): the layer whose patch width is closest to w (DirectPyramidFeatureExtractor.cpp:67-73, ImagePyramid.cpp:307-310)
    olc = int(round(math.log(0.5) / math.log(inc)))
    incr = math.pow(0.5, 1.0 / olc)
    power = math.log(pw / 160.0) / math.log(incr)
    idx = int(math.floor(abs(power) + 0.5)) * (1 if power >= 0 else -1)
    scale, img = by_index[idx]
    x, y = L.fdo_cvround((320 - 80) * scale), L.fdo_cvround((240 - 80) * scale)
    ow, oh = L.fdo_cvround(pw / scale), L.fdo_cvround(ph / scale)
    patch = fo.hq64(img[y:y + ph, x:x + pw])
    assert [int(v) for v in out["BYBOX"][0]] == [L.fdo_cvround(x / scale) + ow // 2, L.fdo_cvround(y / scale) + oh // 2, ow, oh, _fnv(patch)]
    lvl, fout, prob, pos = wo.eval(patch.reshape(1, -1))
    dist, sprob, spos = so.eval(patch.reshape(1, -1))
    c = out["CLASSIFY"][0]
    assert int(c[0]) == int(pos[0]) and abs(float(c[1]) - float(prob[0])) <= 1e-9
    assert int(c[2]) == int(spos[0]) and abs(float(c[3]) - float(sprob[0])) <= 1e-9

    # the `single` psvm detector on a crop: whole image and inside a region of interest (ffpDetectApp.cpp:557,591)
    crop = np.ascontiguousarray(frame[:120, :160])
    kw2 = dict(kw, min_scale_factor=0.2, max_scale_factor=0.5)
    _, layers2 = fo.pyramid(crop, inc, 0.2, 0.5)
    by_index2 = {i: (s, im) for i, s, im in layers2}
    for tag, roi in (("SINGLE_DET", (0, 0, 0, 0)), ("SINGLE_ROI", (30, 20, 100, 90))):
        ref = fo.detect_frame(kw2, None, so, crop, roi=roi)["detections"]
        items = out.get(tag + "_ITEM", [])
        assert int(out[tag][0][0]) == len(items) == len(ref) and len(ref) > 0, tag
        for it, d in zip(items, ref):
            assert [int(v) for v in it[:4]] == [int(d["center_x"]), int(d["center_y"]), int(d["width"]), int(d["height"])], tag
            assert abs(float(it[4]) - float(d["probability"])) <= 1e-9, tag
