"""The oracle's feature spaces (oracle/fd_features.c) against tests/golden/features.npz (cv2 4.13 for
cv::Sobel / cv::equalizeHist / cv::dft; the reference's OWN compiled GradientBinningFilter, HistogramFilter,
SpatialHistogramFilter, HogFilter, ExtendedHogFilter and LbpFilter for the in-repo arithmetic) and, where the
running environment has them, against cv2 and oracle/_ref live.

Tolerance: u8 results and the histogram feature vectors are bit-exact; the whitening chain goes through OpenCV's
float32 FFT, which the oracle evaluates in double - its u8 stage is compared exactly on the golden patches (no
mismatch among them) and the final float vector within 1e-4 (north_star)."""
import ast
import os
import zlib

import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fo(built):
    from oracle import fdoracle
    return fdoracle


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "features.npz"))


@pytest.fixture(scope="module")
def img():
    return np.ascontiguousarray(syn.synthetic_frame(0)[60:180, 200:360])


def test_sobel_matches_cv2_golden(fo, g, img):
    for name, src in (("img", img), ("noise", g["noise"])):
        for k in (1, 3):
            assert np.array_equal(fo.gradient(src, k), g["sobel_%s_k%d" % (name, k)]), (name, k)


def test_binning_luts_match_reference_golden(fo, g):
    for bins, sg in ((9, 0), (18, 1), (8, 1)):
        one, two = fo.gradient_bin_luts(bins, sg)
        crc = g["lut_crc_%d_%d" % (bins, sg)]
        assert zlib.crc32(one.tobytes()) == crc[0] and zlib.crc32(two.tobytes()) == crc[1]
        assert np.array_equal(two[::251], g["lut_sample_%d_%d" % (bins, sg)])


def test_lbp_matches_reference_golden(fo, g, img):
    for t in range(4):
        assert np.array_equal(fo.lbp(img, t), g["lbp_%d" % t]), t


def _patch(g, img, i):
    is_noise, x, y, w, h = g["patch_boxes"][i]
    return np.ascontiguousarray((g["noise"] if is_noise else img)[y:y + h, x:x + w])


def test_equalize_hist_matches_cv2_golden(fo, g, img):
    for i in range(len(g["patch_boxes"])):
        assert np.array_equal(fo.equalize_hist(_patch(g, img, i)), g["histeq_%d" % i]), i
    assert np.array_equal(fo.equalize_hist(np.full((20, 20), 77, np.uint8)), np.full((20, 20), 77, np.uint8))


def test_whitening_chain_matches_cv2_golden(fo, g, img):
    for i in range(len(g["patch_boxes"])):
        p = _patch(g, img, i)
        u8, _ = fo.whitening(p)
        assert np.array_equal(u8, g["whi_u8_%d" % i]), i
        F = fo.Features(syn.feature_desc(kind="whi"), p.shape[1], p.shape[0])
        v = F.patch(p[:, :, None], 0, 0)
        assert np.allclose(v, g["whi_vec_%d" % i], rtol=0, atol=1e-4), i
        assert abs(float(np.linalg.norm(v.astype(np.float64))) - 1.0) < 1e-3


def test_patch_histograms_match_reference_golden(fo, g, img):
    cases = [ast.literal_eval(str(c)) for c in g["hist_cases"]]
    boxes = g["hist_boxes"]
    k = 0
    for ci, kw in enumerate(cases):
        d = syn.feature_desc(**kw)
        for (pw, ph) in ((20, 20), (30, 30), (32, 16)):
            F = fo.Features(d, pw, ph)
            fl = F.filter_layer(img)
            want = g["hist_%d_%dx%d" % (ci, pw, ph)]
            assert F.dim == want.shape[1]
            for i in range(want.shape[0]):
                assert tuple(boxes[k][:3]) == (ci, pw, ph)
                got = F.patch(fl, int(boxes[k][3]), int(boxes[k][4]))
                assert np.array_equal(got.view(np.uint32), want[i].view(np.uint32)), (kw, pw, ph, i)
                k += 1


def test_features_live_against_compiled_reference(fo, img):
    if not fo.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference)")
    rng = np.random.default_rng(5)
    for bins, sg in ((9, 0), (12, 1)):
        a, b = fo.gradient_bin_luts(bins, sg), fo.ref_gradient_bin_luts(bins, sg)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for kw in (dict(kind="hog", cell_size=4, interpolate_cells=True, normalization="l2hys"),
               dict(kind="hog", block_size=3, cell_size=4), dict(kind="ehog", cell_size=4),
               dict(kind="lbp", lbp_type="lbp8uniform", cell_size=8, block_size=2, concatenate=True)):
        d = syn.feature_desc(**kw)
        F = fo.Features(d, 24, 24)
        fl = F.filter_layer(img)
        bins = fo.lib().fdo_lbp_bins(d.lbp_type) if d.kind == capi.FDB_FEATURE_LBP else d.bins
        for _ in range(10):
            x = int(rng.integers(0, img.shape[1] - 24)); y = int(rng.integers(0, img.shape[0] - 24))
            a = F.patch(fl, x, y); b = fo.ref_patch_histogram(d, bins, fl, x, y, 24, 24)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), kw


def test_primitives_live_against_cv2(fo, img):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    rng = np.random.default_rng(11)
    for _ in range(10):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 90))
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        for k, sc in ((1, 0.5), (3, 0.125)):
            gx = cv2.Sobel(src, cv2.CV_8U, 1, 0, ksize=k, scale=sc, delta=127)
            gy = cv2.Sobel(src, cv2.CV_8U, 0, 1, ksize=k, scale=sc, delta=127)
            assert np.array_equal(fo.gradient(src, k), np.stack([gx, gy], axis=-1))
        assert np.array_equal(fo.equalize_hist(src), cv2.equalizeHist(src))


def test_feature_space_detection_modes(fo):
    """five-stage cascade with the SVM in HOG space, and the `single` psvm detector: the whole-frame oracle agrees
    with its own per-window pieces (extract -> svm)."""
    det_kw, wvm, _ = syn.landmark_models("FaceFrontal")
    frame = syn.synthetic_frame(0)
    d = syn.feature_desc(kind="hog")
    F = fo.Features(d, 20, 20)
    r1 = fo.detect_frame(det_kw, fo.Wvm(wvm), None, frame, stage=capi.FDB_STAGE_OE, want_dense=False)
    cand = r1["detections"]
    lxy = np.stack([cand["layer"], cand["x"], cand["y"]], axis=1).astype(np.int32)
    vec = F.extract(det_kw, frame, lxy)
    assert vec.shape == (len(cand), 144) and np.isfinite(vec).all()
    svm_model = syn.make_feature_svm(vec, seed=1, num_sv=64, gamma=0.2)
    so = fo.Svm(svm_model)
    dist, prob, pos = so.eval(vec)
    r2 = fo.detect_frame(det_kw, fo.Wvm(wvm), so, frame, stage=capi.FDB_STAGE_SVM, want_dense=False, svm_features=F)
    got = r2["detections"]
    assert set(got["window"]) == set(cand["window"][pos.astype(bool)])
    for row in got:
        i = int(np.nonzero(cand["window"] == row["window"])[0][0])
        assert row["svm_distance"] == dist[i]
    # single psvm detector over a ROI-free small frame
    small = np.ascontiguousarray(frame[:160, :200])
    kw = dict(det_kw, min_scale_factor=0.2, max_scale_factor=0.3)
    r3 = fo.detect_frame(kw, None, so, small, want_dense=True, svm_features=F)
    assert r3["svm_dense"].shape == (r3["windows"],) and r3["windows"] > 0
    assert np.array_equal(r3["detections"]["window"], np.nonzero(r3["svm_dense"] >= 0)[0])
