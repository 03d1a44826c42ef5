"""GPU parity of the feature spaces (features.cu through the C ABI) against the CPU oracle and the committed
golden vectors (tests/golden/features.npz: cv2 4.13 + the reference's own compiled filters).

Bar: u8 feature vectors (gray, histeq) bit-exact; float32 histogram features (hog / ehog / lbp) bit-exact against the
oracle (same operation order); whi within 1e-4 with the fraction of windows whose whitened u8 patch differs printed
(the transform is OpenCV's float32 FFT in the reference, double here and in the oracle); SVM distances within 1e-4;
detections identical."""
import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn
from featuredetection_b200.detector import SlidingWindowCascade

pytestmark = pytest.mark.gpu
TOL = 1e-4

CASES = [
    dict(kind="gray"), dict(kind="histeq"), dict(kind="whi"),
    dict(kind="hog"), dict(kind="hog", interpolate_cells=True), dict(kind="hog", interpolate_bins=False, gradient_kernel=3),
    dict(kind="hog", cell_size=6, normalization="l2hys"), dict(kind="hog", normalization="l1sqrt", interpolate_cells=True),
    dict(kind="hog", block_size=2), dict(kind="hog", block_size=2, signed_and_unsigned=True, bins=8, signed_gradients=True),
    dict(kind="ehog"), dict(kind="ehog", signed_and_unsigned=True, bins=18, signed_gradients=True, interpolate_cells=True),
    dict(kind="lbp", cell_size=10), dict(kind="lbp", lbp_type="lbp8", cell_size=10, interpolate_cells=True),
    dict(kind="lbp", lbp_type="lbp4", block_size=2, concatenate=True, normalization="l2hys"),
    dict(kind="lbp", lbp_type="lbp4rotated", block_size=2, concatenate=False, normalization="l1norm"),
]


def _oracle():
    from oracle import fdoracle as fo
    return fo


def _windows(casc, rng, n):
    out = []
    for L in casc.layers():
        for _ in range(n):
            out.append((L["index"], int(rng.integers(0, L["width"] - casc.patch[0] + 1)),
                        int(rng.integers(0, L["height"] - casc.patch[1] + 1))))
    # corners: layer borders exercise the reflect / replicate handling of the layer filters
    L = casc.layers()[0]
    out += [(L["index"], 0, 0), (L["index"], L["width"] - casc.patch[0], L["height"] - casc.patch[1])]
    return np.array(out, np.int32)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_feature_vectors_match_oracle(ctx, face_models, case):
    fo = _oracle()
    det_kw, wvm, _ = face_models
    kw = CASES[case]
    d = syn.feature_desc(**kw)
    casc = SlidingWindowCascade(ctx, det_kw, wvm, None, feature=d)
    casc.prepare(640, 480, 1)
    F = fo.Features(d, 20, 20)
    assert (casc.feature_dim, casc.feature_dtype) == (F.dim, F.dtype)
    frame = syn.synthetic_frame(5)
    lxy = _windows(casc, np.random.default_rng(case), 6)
    got = casc.extract_features(frame, lxy)
    want = F.extract(det_kw, frame, lxy)
    if kw["kind"] == "whi":
        assert np.allclose(got, want, rtol=0, atol=TOL)
    elif F.is_float:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "max diff %g" % np.abs(got - want).max()
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("patch", [(24, 24), (32, 16), (16, 24), (32, 24)])
def test_feature_vectors_other_patch_sizes(ctx, patch):
    fo = _oracle()
    name = {(24, 24): "LeftLipCorner", (32, 16): "RightEyeCenter", (16, 24): "LeftEarCenter", (32, 24): "NoseTip"}[patch]
    det_kw, wvm, _ = syn.landmark_models(name)
    frame = np.ascontiguousarray(syn.synthetic_frame(6)[:120, :160])
    for kw in (dict(kind="whi"), dict(kind="hog", cell_size=4, interpolate_cells=True), dict(kind="lbp", cell_size=8), dict(kind="histeq")):
        d = syn.feature_desc(**kw)
        casc = SlidingWindowCascade(ctx, det_kw, wvm, None, feature=d)
        casc.prepare(160, 120, 1)
        F = fo.Features(d, *patch)
        lxy = _windows(casc, np.random.default_rng(3), 8)
        got = casc.extract_features(frame, lxy)
        want = F.extract(det_kw, frame, lxy)
        if kw["kind"] == "whi":
            assert np.allclose(got, want, rtol=0, atol=TOL)
        elif F.is_float:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        else:
            assert np.array_equal(got, want)


def test_feature_vectors_match_golden(ctx, face_models):
    """the CUDA path against the committed vectors of cv2 / the compiled reference, without the oracle in between:
    a 160x120 crop is scanned at scale 1 (layer 0 = the frame itself), so layer filters see the golden image"""
    import ast
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "features.npz"))
    img = np.ascontiguousarray(syn.synthetic_frame(0)[60:180, 200:360])
    _, wvm, _ = face_models
    cases = [ast.literal_eval(str(c)) for c in g["hist_cases"]]
    boxes = g["hist_boxes"]
    k = 0
    for ci, kw in enumerate(cases):
        for (pw, ph) in ((20, 20), (30, 30), (32, 16)):
            want = g["hist_%d_%dx%d" % (ci, pw, ph)]
            rows = boxes[k:k + want.shape[0]]
            k += want.shape[0]
            if (pw, ph) != (20, 20):
                continue  # detector patch size = the WVM's 20x20 here
            det_kw = dict(incremental_scale_factor=0.9, min_scale_factor=0.95, max_scale_factor=1.0, patch_width=pw, patch_height=ph,
                          step_x=1, step_y=1, oe_dist=5.0, oe_ratio=0.0)
            casc = SlidingWindowCascade(ctx, det_kw, wvm, None, feature=syn.feature_desc(**kw))
            casc.prepare(160, 120, 1)
            assert [L["index"] for L in casc.layers()] == [0]
            lxy = np.array([(0, r[3], r[4]) for r in rows], np.int32)
            got = casc.extract_features(img, lxy)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), kw
    # whi / histeq patches of the golden file that come from the image crop (not the noise image)
    det_kw = dict(incremental_scale_factor=0.9, min_scale_factor=0.95, max_scale_factor=1.0, patch_width=20, patch_height=20,
                  step_x=1, step_y=1, oe_dist=5.0, oe_ratio=0.0)
    sel = [i for i, b in enumerate(g["patch_boxes"]) if not b[0] and (b[3], b[4]) == (20, 20)]
    lxy = np.array([(0, g["patch_boxes"][i][1], g["patch_boxes"][i][2]) for i in sel], np.int32)
    casc = SlidingWindowCascade(ctx, det_kw, wvm, None, feature=syn.feature_desc(kind="histeq"))
    casc.prepare(160, 120, 1)
    got = casc.extract_features(img, lxy)
    for j, i in enumerate(sel):
        assert np.array_equal(got[j].reshape(20, 20), g["histeq_%d" % i])
    casc = SlidingWindowCascade(ctx, det_kw, wvm, None, feature=syn.feature_desc(kind="whi"))
    casc.prepare(160, 120, 1)
    got = casc.extract_features(img, lxy)
    for j, i in enumerate(sel):
        assert np.allclose(got[j], g["whi_vec_%d" % i], rtol=0, atol=TOL)


def _hog_svm(fo, det_kw, wvm, kw, frames, num_sv=256):
    d = syn.feature_desc(**kw)
    F = fo.Features(d, det_kw["patch_width"], det_kw["patch_height"])
    r = fo.detect_frame(det_kw, fo.Wvm(wvm), None, frames[0], stage=capi.FDB_STAGE_WVM, want_dense=False)
    cand = r["detections"]
    vec = F.extract(det_kw, frames[0], np.stack([cand["layer"], cand["x"], cand["y"]], axis=1))
    svm = syn.make_feature_svm(vec, seed=7, num_sv=num_sv, gamma=syn.FEATURE_GAMMA[kw["kind"]])
    # threshold at the median distance of these candidates, so that the stage keeps about half of them
    dist, _, _ = fo.Svm(svm).eval(vec)
    svm.threshold = float(np.float32(np.median(dist)))
    return d, F, svm


@pytest.mark.parametrize("kind", ["hog", "whi", "lbp", "histeq"])
def test_five_stage_with_feature_space_svm(ctx, face_models, kind):
    """BASELINE configs[1]: FaceFrontal WVM -> SVM cascade with the SVM in HOG (WHI, LBP, histeq) space"""
    fo = _oracle()
    det_kw, wvm, _ = face_models
    frames = syn.synthetic_frames(20, 3)
    kw = dict(kind=kind, cell_size=10) if kind == "lbp" else dict(kind=kind)
    d, F, svm = _hog_svm(fo, det_kw, wvm, kw, frames)
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm, feature=d)
    casc.prepare(640, 480, 3)
    for stage in (capi.FDB_STAGE_SVM, capi.FDB_STAGE_NMS):
        dets = casc.detect(frames, stage=stage)
        wo, so = fo.Wvm(wvm), fo.Svm(svm)
        for k in range(3):
            ref = fo.detect_frame(det_kw, wo, so, frames[k], stage=stage, frame_index=k, want_dense=False, svm_features=F)["detections"]
            mine = dets[dets["frame"] == k]
            assert list(mine["window"]) == list(ref["window"]), (kind, stage, k)
            assert np.allclose(mine["svm_distance"], ref["svm_distance"], rtol=0, atol=TOL)
            assert np.allclose(mine["svm_probability"], ref["svm_probability"], rtol=0, atol=TOL)
    assert len(dets) > 0


@pytest.mark.parametrize("kind", ["hq64", "hog", "whi"])
def test_single_psvm_detector_all_windows(ctx, face_models, kind):
    """ffpDetectApp `single` detector with classifier psvm (BASELINE configs[3] "RBF-SVM"): every window classified"""
    fo = _oracle()
    det_kw, wvm, svm_u8 = face_models
    kw = dict(det_kw, min_scale_factor=0.09, max_scale_factor=0.16)
    frames = np.ascontiguousarray(syn.synthetic_frames(30, 2)[:, :240, :320])
    if kind == "hq64":
        d, F, svm = None, None, syn.make_svm(20, 20, seed=5, num_sv=128)
    else:
        d = syn.feature_desc(kind=kind)
        F = fo.Features(d, 20, 20)
        _, layers = fo.pyramid(frames[0], kw["incremental_scale_factor"], kw["min_scale_factor"], kw["max_scale_factor"])
        rng = np.random.default_rng(12)
        lxy = [(idx, int(rng.integers(0, img.shape[1] - 20)), int(rng.integers(0, img.shape[0] - 20))) for idx, _, img in layers for _ in range(40)]
        vec = F.extract(kw, frames[0], np.array(lxy, np.int32))
        svm = syn.make_feature_svm(vec, seed=7, num_sv=128, gamma=syn.FEATURE_GAMMA[kind])
        svm.threshold = float(np.float32(np.median(fo.Svm(svm).eval(vec)[0])))
    casc = SlidingWindowCascade(ctx, kw, None, svm, feature=d)
    casc.prepare(320, 240, 2)
    dets, dist = casc.detect_single(frames)
    so = fo.Svm(svm)
    for k in range(2):
        ref = fo.detect_frame(kw, None, so, frames[k], frame_index=k, svm_features=F)
        assert dist.shape[1] == ref["windows"]
        assert np.allclose(dist[k], ref["svm_dense"], rtol=0, atol=TOL), np.abs(dist[k] - ref["svm_dense"]).max()
        mine = dets[dets["frame"] == k]
        assert list(mine["window"]) == list(ref["detections"]["window"])
        assert np.allclose(mine["probability"], ref["detections"]["probability"], rtol=0, atol=TOL)


def test_whitening_u8_mismatch_fraction(ctx, face_models):
    """prints how many windows' whi vectors differ visibly from the oracle (a whitened pixel on the other side of a
    rounding boundary shows as a difference > 1e-3 after equalisation and normalisation)"""
    fo = _oracle()
    det_kw, wvm, _ = face_models
    d = syn.feature_desc(kind="whi")
    casc = SlidingWindowCascade(ctx, det_kw, wvm, None, feature=d)
    casc.prepare(640, 480, 1)
    F = fo.Features(d, 20, 20)
    frame = syn.synthetic_frame(9)
    rng = np.random.default_rng(0)
    lxy = _windows(casc, rng, 40)
    got = casc.extract_features(frame, lxy)
    want = F.extract(det_kw, frame, lxy)
    bad = int((np.abs(got - want).max(axis=1) > 1e-3).sum())
    print("whi: %d of %d windows differ from the oracle by more than 1e-3" % (bad, len(lxy)))
    assert bad == 0


def test_feature_errors(ctx, face_models):
    det_kw, wvm, svm = face_models
    lib = ctx.lib
    with pytest.raises(capi.FdbError):  # u8 hq64 SVM cannot classify float HOG vectors
        c = SlidingWindowCascade(ctx, det_kw, wvm, svm, feature=syn.feature_desc(kind="hog"))
        c.prepare(640, 480, 1)
    with pytest.raises(capi.FdbError):
        SlidingWindowCascade(ctx, det_kw, wvm, None, feature=syn.feature_desc(kind="hog", cell_size=0))
    with pytest.raises(capi.FdbError):
        SlidingWindowCascade(ctx, det_kw, wvm, None, feature=syn.feature_desc(kind="hog", signed_and_unsigned=True, bins=9))
    c = SlidingWindowCascade(ctx, det_kw, wvm, None, feature=syn.feature_desc(kind="hog"))
    c.prepare(640, 480, 1)
    with pytest.raises(capi.FdbError):  # window outside its layer: the reference returns an empty patch
        c.extract_features(syn.synthetic_frame(0), np.array([[22, 90, 60]], np.int32))
    with pytest.raises(capi.FdbError):
        c.extract_features(syn.synthetic_frame(0), np.array([[99, 0, 0]], np.int32))
