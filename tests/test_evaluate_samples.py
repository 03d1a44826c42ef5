"""condensation::WvmSvmModel::evaluate(image, samples) (WvmSvmModel.cpp:74-119) - the tracker's sparse use of the two
classifiers: oracle properties on the CPU, fdb_evaluate_samples against the oracle on the GPU."""
import numpy as np
import pytest

from featuredetection_b200 import synthetic as syn


def _samples(rng, n, W=640, H=480):
    """particles like the tracker's: centres anywhere (some outside), widths around the pyramid's range (some outside)"""
    w = rng.integers(100, 460, n)
    s = np.stack([rng.integers(-20, W + 20, n), rng.integers(-20, H + 20, n), w, w], axis=1).astype(np.int32)
    s[: n // 8] = s[n // 8: 2 * (n // 8)][: n // 8]          # duplicates -> same patch, classified once
    s[n // 8: n // 8 + 6, :2] += 1                           # near-duplicates that round to the same layer pixel
    return s


def test_oracle_evaluate_samples_properties(built, face_models):
    from oracle import fdoracle as fo
    det_kw, wvm, svm = face_models
    frame = syn.synthetic_frame(5)
    rng = np.random.default_rng(1)
    smp = _samples(rng, 120)
    wo, so = fo.Wvm(wvm), fo.Svm(svm)
    target, weight = fo.evaluate_samples(det_kw, wo, so, frame, smp)
    assert target.shape == (120,) and weight.shape == (120,)
    assert np.all(weight >= 0) and np.all(weight <= 1)
    assert np.all(weight[target] > 0)
    # samples without a patch: width outside the pyramid or window outside the layer
    off = (smp[:, 2] < 120) | (smp[:, 2] > 420) | (smp[:, 0] < 0) | (smp[:, 1] < 0)
    assert np.all(weight[off & (smp[:, 2] < 110)] == 0)
    # equal samples get equal results; the SVM is applied to at most 8 distinct patches
    assert np.array_equal(weight[:15], weight[15:30]) and np.array_equal(target[:15], target[15:30])
    t_all, w_all = fo.evaluate_samples(det_kw, wo, so, frame, smp, max_svm_patches=0)
    assert np.all(w_all[~t_all & (w_all > 0)] <= 0.5 + 1e-12) and t_all.sum() >= target.sum()
    # no SVM: weights are 0.5 * P_wvm, no targets
    t0, w0 = fo.evaluate_samples(det_kw, wo, None, frame, smp)
    assert not t0.any() and np.all(w0 <= 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("top", [8, 0, 3])
def test_evaluate_samples_matches_oracle(ctx, face_models, face_models_noexit, top):
    from oracle import fdoracle as fo
    from featuredetection_b200.detector import SlidingWindowCascade
    for models, seed in ((face_models, 2), (face_models_noexit, 3)):
        det_kw, wvm, svm = models
        casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
        casc.prepare(640, 480, 1)
        wo, so = fo.Wvm(wvm), fo.Svm(svm)
        rng = np.random.default_rng(seed)
        frame = syn.synthetic_frame(seed)
        smp = _samples(rng, 400)
        target, weight = casc.evaluate_samples(frame, smp, max_svm_patches=top)
        rt, rw = fo.evaluate_samples(det_kw, wo, so, frame, smp, max_svm_patches=top)
        assert np.array_equal(target, rt), (top, int(target.sum()), int(rt.sum()))
        assert np.max(np.abs(weight - rw)) <= 1e-9, np.max(np.abs(weight - rw))
        assert (weight > 0).sum() > 50


@pytest.mark.gpu
def test_evaluate_samples_edge_cases(ctx, face_models):
    from featuredetection_b200 import capi
    from featuredetection_b200.detector import SlidingWindowCascade
    det_kw, wvm, svm = face_models
    casc = SlidingWindowCascade(ctx, det_kw, wvm, None)   # no second stage
    casc.prepare(640, 480, 1)
    frame = syn.synthetic_frame(0)
    t, w = casc.evaluate_samples(frame, np.zeros((0, 4), np.int32))
    assert len(t) == 0 and len(w) == 0
    t, w = casc.evaluate_samples(frame, np.array([[320, 240, 5, 5], [320, 240, 0, 0], [-500, 240, 200, 200], [320, 240, 200, 200]], np.int32))
    assert not t.any() and np.all(w[:3] == 0) and 0 < w[3] <= 0.5
    single = SlidingWindowCascade(ctx, det_kw, None, svm)
    single.prepare(640, 480, 1)
    with pytest.raises(capi.FdbError):
        single.evaluate_samples(frame, np.array([[320, 240, 200, 200]], np.int32))


@pytest.mark.gpu
def test_face_then_feature_detectors_in_the_face_box(ctx, face_models):
    """ffpDetectApp.cpp:553-596: face detector, then the landmark detectors inside the bounds of the first face patch"""
    from oracle import fdoracle as fo
    from featuredetection_b200 import capi
    from featuredetection_b200.detector import SlidingWindowCascade, detect_face_features
    det_kw, wvm, svm = face_models
    face = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    face.prepare(640, 480, 1)
    names = ["RightEyeCenter", "NoseTip"]
    models = [syn.landmark_models(n) for n in names]
    feats = []
    for kw, w, s in models:
        c = SlidingWindowCascade(ctx, kw, w, s)
        c.prepare(640, 480, 1)
        feats.append(c)
    checked = 0
    for k in (0, 1, 2):
        frame = syn.synthetic_frame(k)
        fd, per = detect_face_features(face, feats, frame)
        ref = fo.detect_frame(det_kw, fo.Wvm(wvm), fo.Svm(svm), frame, stage=capi.FDB_STAGE_NMS)["detections"]
        assert list(fd["window"]) == list(ref["window"])
        if len(ref) == 0:
            assert all(len(p) == 0 for p in per)
            continue
        f = ref[0]
        roi = (int(f["center_x"] - f["width"] // 2), int(f["center_y"] - f["height"] // 2), int(f["width"]), int(f["height"]))
        for (kw, w, s), mine in zip(models, per):
            r = fo.detect_frame(kw, fo.Wvm(w), fo.Svm(s), frame, stage=capi.FDB_STAGE_NMS, roi=roi)["detections"]
            assert list(mine["window"]) == list(r["window"]) and np.array_equal(mine["center_x"], r["center_x"])
            assert np.allclose(mine["probability"], r["probability"], rtol=0, atol=1e-9)
            checked += 1
    assert checked > 0
