"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs. Bit-exact for pyramid pixels, equalised patches, window indices and cascade
levels; scores within 1e-4 (north_star tolerance)."""
import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn
from featuredetection_b200.detector import SlidingWindowCascade, ProbabilisticWvmClassifier, ProbabilisticSvmClassifier

pytestmark = pytest.mark.gpu
TOL = 1e-4  # north_star: per-patch scores within 1e-4, indices bit-exact


def _oracle():
    from oracle import fdoracle as fo
    return fo


def test_pyramid_layers_bit_exact(ctx, face_models):
    fo = _oracle()
    det_kw, wvm, svm = face_models
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(640, 480, 2)
    for k in (0, 3):
        frame = syn.synthetic_frame(k)
        olc, layers = fo.pyramid(frame, det_kw["incremental_scale_factor"], det_kw["min_scale_factor"], det_kw["max_scale_factor"])
        infos = casc.layers()
        assert [L["index"] for L in infos] == [i for i, _, _ in layers]
        for (idx, scale, img), info in zip(layers, infos):
            got = casc.pyramid_layer(frame, idx)
            assert got.shape == img.shape
            assert np.array_equal(got, img), "layer %d differs" % idx
            assert info["scale"] == scale


@pytest.mark.parametrize("size", [(641, 479), (333, 250), (97, 61)])
def test_pyramid_odd_sizes(ctx, face_models, size):
    fo = _oracle()
    det_kw, wvm, svm = face_models
    kw = dict(det_kw, min_scale_factor=0.11, max_scale_factor=0.6)
    W, H = size
    casc = SlidingWindowCascade(ctx, kw, wvm, svm)
    casc.prepare(W, H, 1)
    frame = syn.synthetic_frame(11, W, H)
    _, layers = fo.pyramid(frame, kw["incremental_scale_factor"], kw["min_scale_factor"], kw["max_scale_factor"])
    assert len(layers) == len(casc.layers())
    for idx, scale, img in layers:
        assert np.array_equal(casc.pyramid_layer(frame, idx), img), "layer %d differs" % idx


def test_extract_patches_bit_exact(ctx, face_models):
    fo = _oracle()
    det_kw, wvm, svm = face_models
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(640, 480, 1)
    frame = syn.synthetic_frame(2)
    ref = fo.detect_frame(det_kw, fo.Wvm(wvm), fo.Svm(svm), frame, stage=capi.FDB_STAGE_WVM, want_patches=True)
    got = casc.extract_patches(frame)
    assert got.shape == ref["patches"].shape == (16185, 400)
    assert np.array_equal(got, ref["patches"])


@pytest.mark.parametrize("profile", ["realistic", "no-exit"])
def test_dense_scores_and_levels(ctx, face_models, face_models_noexit, profile):
    fo = _oracle()
    det_kw, wvm, svm = face_models if profile == "realistic" else face_models_noexit
    if profile == "no-exit":  # every window is a stage-1 positive
        det_kw = dict(det_kw, max_positives_per_frame=20000)
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    nframes = 3 if profile == "realistic" else 1
    casc.prepare(640, 480, nframes)
    frames = syn.synthetic_frames(20, nframes)
    dets, dense = casc.detect(frames, stage=capi.FDB_STAGE_WVM, want_dense=True, det_cap=1 << 17)
    wo = fo.Wvm(wvm)
    for k in range(nframes):
        okw = {k2: v for k2, v in det_kw.items() if k2 != "max_positives_per_frame"}
        ref = fo.detect_frame(okw, wo, None, frames[k], stage=capi.FDB_STAGE_WVM, det_cap=1 << 17)
        assert ref["windows"] == dense.shape[1] == 16185
        assert np.array_equal(dense[k]["level"], ref["dense"]["level"])
        assert np.max(np.abs(dense[k]["fout"] - ref["dense"]["fout"])) <= TOL
        mine = dets[dets["frame"] == k]
        assert list(mine["window"]) == list(ref["detections"]["window"])
        for f in ("layer", "x", "y", "center_x", "center_y", "width", "height", "wvm_level"):
            assert np.array_equal(mine[f], ref["detections"][f]), f
        assert np.allclose(mine["wvm_probability"], ref["detections"]["wvm_probability"], rtol=0, atol=TOL)


@pytest.mark.parametrize("stage", [capi.FDB_STAGE_OE, capi.FDB_STAGE_SVM, capi.FDB_STAGE_NMS])
def test_five_stage_detections(ctx, face_models, stage):
    fo = _oracle()
    det_kw, wvm, svm = face_models
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(640, 480, 2)  # 5 frames through a 2-frame batch: exercises chunking
    frames = syn.synthetic_frames(40, 5)
    dets = casc.detect(frames, stage=stage)
    wo, so = fo.Wvm(wvm), fo.Svm(svm)
    total = 0
    for k in range(5):
        ref = fo.detect_frame(det_kw, wo, so, frames[k], stage=stage, frame_index=k)["detections"]
        mine = dets[dets["frame"] == k]
        assert list(mine["window"]) == list(ref["window"]), "frame %d stage %d" % (k, stage)
        assert np.array_equal(mine["probability"], ref["probability"])
        if stage >= capi.FDB_STAGE_SVM:
            assert np.allclose(mine["svm_distance"], ref["svm_distance"], rtol=0, atol=TOL)
            assert np.allclose(mine["svm_probability"], ref["svm_probability"], rtol=0, atol=TOL)
        total += len(ref)
    assert total == len(dets)


def test_roi_detection(ctx, face_models):
    fo = _oracle()
    det_kw, wvm, svm = face_models
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(640, 480, 1)
    frame = syn.synthetic_frame(7)
    wo, so = fo.Wvm(wvm), fo.Svm(svm)
    for roi in [(100, 60, 400, 350), (-20, -10, 700, 300), (0, 0, 640, 480)]:
        for stage in (capi.FDB_STAGE_WVM, capi.FDB_STAGE_SVM):
            ref = fo.detect_frame(det_kw, wo, so, frame, stage=stage, roi=roi)["detections"]
            mine = casc.detect_roi(frame, roi, stage=stage)
            assert list(mine["window"]) == list(ref["window"]), (roi, stage)
            assert np.array_equal(mine["center_x"], ref["center_x"]) and np.array_equal(mine["center_y"], ref["center_y"])


def test_classifier_entry_points(ctx, face_models):
    """ProbabilisticClassifier::getProbability over given feature vectors."""
    fo = _oracle()
    det_kw, wvm, svm = face_models
    rng = np.random.default_rng(5)
    frame = syn.synthetic_frame(9)
    _, layers = fo.pyramid(frame, det_kw["incremental_scale_factor"], det_kw["min_scale_factor"], det_kw["max_scale_factor"])
    img = layers[0][2]
    patches = np.stack([fo.hq64(img[y:y + 20, x:x + 20]).ravel()
                        for y, x in zip(rng.integers(0, 50, 300), rng.integers(0, 70, 300))])
    patches[0] = 0; patches[1] = 255; patches[2] = rng.integers(0, 256, 400)  # degenerate inputs
    gw = ProbabilisticWvmClassifier(ctx, wvm)
    level, fout, prob, pos = gw.get_probability(patches)
    rl, rf, rp, rpos = fo.Wvm(wvm).eval(patches)
    assert np.array_equal(level, rl) and np.array_equal(pos, rpos)
    assert np.max(np.abs(fout - rf)) <= TOL and np.max(np.abs(prob - rp)) <= TOL
    gs = ProbabilisticSvmClassifier(ctx, svm)
    dist, sp, spos = gs.get_probability(patches)
    rd, rsp, rspos = fo.Svm(svm).eval(patches)
    assert np.max(np.abs(dist - rd)) <= TOL and np.max(np.abs(sp - rsp)) <= TOL
    assert np.array_equal(spos, rspos)
    # float32 support vectors (RbfKernel::computeSumOfSquaredDifferences_any<float>)
    fsv = rng.normal(0, 1, (64, 144)).astype(np.float32)
    fmodel = syn.SvmModel(fsv, rng.normal(0, 1, 64).astype(np.float32), gamma=0.2, bias=0.1, threshold=0.05)
    x = rng.normal(0, 1, (40, 144)).astype(np.float32)
    d2, p2, q2 = ProbabilisticSvmClassifier(ctx, fmodel).get_probability(x)
    r2, rp2, rq2 = fo.Svm(fmodel).eval(x)
    assert np.max(np.abs(d2 - r2)) <= TOL and np.array_equal(q2, rq2)


def test_limit_reliability_and_empty(ctx, face_models):
    fo = _oracle()
    det_kw, wvm, svm = face_models
    gw = ProbabilisticWvmClassifier(ctx, wvm)
    lv, fo_, pr, pos = gw.get_probability(np.zeros((0, 400), np.uint8))
    assert len(lv) == 0
    gw.set_limit_reliability_filter(0.05)
    import copy
    m2 = copy.copy(wvm); m2.limit_reliability_filter = 0.05
    rng = np.random.default_rng(3)
    patches = rng.integers(0, 256, (200, 400), dtype=np.uint8)
    level, fout, _, _ = gw.get_probability(patches)
    rl, rf, _, _ = fo.Wvm(m2).eval(patches)
    assert np.array_equal(level, rl) and np.max(np.abs(fout - rf)) <= TOL


def test_batch_invariance_full_size(ctx, face_models):
    """Full-size (256-frame) property: results do not depend on batch position or chunking,
    duplicated frames give identical records, and frame k of the batch equals a solo run."""
    det_kw, wvm, svm = face_models
    base = syn.synthetic_frames(100, 8)
    frames = np.concatenate([base] * 32)  # 256 frames
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(640, 480, 256)
    dets, dense = casc.detect(frames, stage=capi.FDB_STAGE_NMS, want_dense=True)
    for r in range(1, 32):
        assert np.array_equal(dense[r * 8:(r + 1) * 8], dense[:8])
    casc2 = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc2.prepare(640, 480, 3)
    dets2, dense2 = casc2.detect(base, stage=capi.FDB_STAGE_NMS, want_dense=True)
    assert np.array_equal(dense2, dense[:8])
    first = dets[dets["frame"] < 8]
    assert np.array_equal(first["window"], dets2["window"]) and np.array_equal(first["frame"], dets2["frame"])
    counts = casc.last_counts()
    assert counts[0] == 256 * 16185 and counts[4] == len(dets)


def test_cuda_path_matches_reference_golden(ctx):
    """The CUDA path against golden vectors produced by the reference's OWN compiled sources
    (tests/golden/ref_classifiers.npz, made by tests/golden/make_ref_golden.py from oracle/_ref)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_classifiers.npz"))
    det_kw, wvm, svm = syn.landmark_models("FaceFrontal")
    _, wvm_ne, _ = syn.landmark_models("FaceFrontal", "no-exit")
    for tag, model in (("realistic", wvm), ("noexit", wvm_ne)):
        lv, fout, pr, pos = ProbabilisticWvmClassifier(ctx, model).get_probability(g["patches"])
        assert np.array_equal(lv, g["wvm_%s_level" % tag]) and np.array_equal(pos, g["wvm_%s_pos" % tag])
        assert np.max(np.abs(fout - g["wvm_%s_fout" % tag])) <= TOL and np.max(np.abs(pr - g["wvm_%s_prob" % tag])) <= TOL
    d, p, q = ProbabilisticSvmClassifier(ctx, svm).get_probability(g["patches"])
    assert np.max(np.abs(d - g["svm_dist"])) <= TOL and np.max(np.abs(p - g["svm_prob"])) <= TOL and np.array_equal(q, g["svm_pos"])
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(640, 480, 2)
    frames = syn.synthetic_frames(0, 2)
    dets, dense = casc.detect(frames, stage=capi.FDB_STAGE_WVM, want_dense=True)
    for k in (0, 1):
        assert np.array_equal(dense[k]["level"], g["frame%d_dense_level" % k])
        assert np.max(np.abs(dense[k]["fout"] - g["frame%d_dense_fout" % k])) <= TOL
    for stage, name in ((capi.FDB_STAGE_WVM, "wvm"), (capi.FDB_STAGE_OE, "oe"), (capi.FDB_STAGE_SVM, "svm"), (capi.FDB_STAGE_NMS, "nms")):
        d = casc.detect(frames, stage=stage)
        for k in (0, 1):
            assert list(d[d["frame"] == k]["window"]) == list(g["frame%d_%s_windows" % (k, name)]), (k, name)
    # extraction entry point: hq64 patches of the golden raw crops are covered by test_extract_patches_bit_exact


def test_fifteen_landmark_geometries(ctx):
    """Every ffpDetectApp cfg geometry (5 patch sizes incl. the generic fallbacks) on one frame: no-exit
    synthetic models with few filters keep it fast; dense records against the oracle."""
    fo = _oracle()
    frame = syn.synthetic_frame(33)
    seen = set()
    for (nm, pw, ph, inc, mn, mx, per, lev, r) in syn.LANDMARK_CONFIGS:
        if (pw, ph, inc, mn, mx) in seen:
            continue
        seen.add((pw, ph, inc, mn, mx))
        wvm = syn.make_wvm(pw, ph, 4, 3, r, seed=900 + len(seen))
        thr = np.full(12, -np.inf, np.float32); thr[::2] = 0.0  # mixed exits
        wvm = wvm.with_thresholds(thr)
        kw = dict(incremental_scale_factor=float(np.float32(inc)), min_scale_factor=float(np.float32(mn)),
                  max_scale_factor=float(np.float32(mx)), patch_width=pw, patch_height=ph, step_x=1, step_y=1,
                  max_positives_per_frame=400000)
        casc = SlidingWindowCascade(ctx, kw, wvm, None)
        casc.prepare(640, 480, 1)
        dets, dense = casc.detect(frame[None], stage=capi.FDB_STAGE_WVM, want_dense=True, det_cap=400000)
        okw = {k: v for k, v in kw.items() if k != "max_positives_per_frame"}
        ref = fo.detect_frame(okw, fo.Wvm(wvm), None, frame, stage=capi.FDB_STAGE_WVM, det_cap=400000)
        assert ref["windows"] == dense.shape[1], nm
        assert np.array_equal(dense[0]["level"], ref["dense"]["level"]), nm
        assert np.max(np.abs(dense[0]["fout"] - ref["dense"]["fout"])) <= TOL, nm
        assert list(dets["window"]) == list(ref["detections"]["window"]), nm


@pytest.mark.parametrize("name", ["LeftEyeCenter", "NoseTip", "LeftEarCenter", "CenterLipUpperOuter", "FaceLeftProfile"])
def test_landmark_models_realistic(ctx, name):
    """Other ffpDetectApp landmark detectors (patch sizes 32x16, 32x24, 16x24, 24x24, 20x20) with their
    calibrated thresholds, whole five-stage cascade on one frame against the oracle."""
    fo = _oracle()
    det_kw, wvm, svm = syn.landmark_models(name)
    frame = syn.synthetic_frame(51)
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(640, 480, 1)
    dets, dense = casc.detect(frame[None], stage=capi.FDB_STAGE_NMS, want_dense=True)
    ref = fo.detect_frame(det_kw, fo.Wvm(wvm), fo.Svm(svm), frame, stage=capi.FDB_STAGE_NMS)
    assert ref["windows"] == dense.shape[1]
    assert np.array_equal(dense[0]["level"], ref["dense"]["level"])
    assert np.max(np.abs(dense[0]["fout"] - ref["dense"]["fout"])) <= TOL
    assert list(dets["window"]) == list(ref["detections"]["window"])
    assert np.allclose(dets["svm_distance"], ref["detections"]["svm_distance"], rtol=0, atol=TOL)


@pytest.mark.parametrize("steps", [(2, 3), (5, 1)])
def test_window_steps_generic_path(ctx, face_models, steps):
    """SlidingWindowDetector(classifier, extractor, stepX, stepY) with steps != 1 (generic kernels)."""
    fo = _oracle()
    det_kw, wvm, svm = face_models
    kw = dict(det_kw, step_x=steps[0], step_y=steps[1])
    casc = SlidingWindowCascade(ctx, kw, wvm, svm)
    casc.prepare(640, 480, 2)
    frames = syn.synthetic_frames(70, 2)
    dets, dense = casc.detect(frames, stage=capi.FDB_STAGE_NMS, want_dense=True)
    wo, so = fo.Wvm(wvm), fo.Svm(svm)
    for k in range(2):
        ref = fo.detect_frame(kw, wo, so, frames[k], stage=capi.FDB_STAGE_NMS, frame_index=k)
        assert ref["windows"] == dense.shape[1]
        assert np.array_equal(dense[k]["level"], ref["dense"]["level"])
        assert np.max(np.abs(dense[k]["fout"] - ref["dense"]["fout"])) <= TOL
        assert list(dets[dets["frame"] == k]["window"]) == list(ref["detections"]["window"])
        assert len(np.unique(ref["dense"]["level"])) >= 4 and len(ref["detections"]) > 0


def test_pitch_partial_batches_and_empty(ctx, face_models):
    """Row pitch > width, batch sizes that do not divide the prepared batch, and an empty batch."""
    import ctypes as C
    from featuredetection_b200.detector import DETECTION_DTYPE
    det_kw, wvm, svm = face_models
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(640, 480, 20)  # chunk = 5 frames, 3 slots
    frames = syn.synthetic_frames(80, 13)
    want = casc.detect(frames, stage=capi.FDB_STAGE_NMS)
    padded = np.zeros((13, 480, 704), np.uint8)
    padded[:, :, :640] = frames
    dets = np.zeros(4096, DETECTION_DTYPE)
    cnt = C.c_int64()
    capi.check(ctx.lib, ctx.lib.fdb_detect_batch(casc.h, padded.ctypes.data, 704, 13, capi.FDB_STAGE_NMS, None,
                                                 dets.ctypes.data, 4096, C.byref(cnt)))
    got = dets[:cnt.value]
    assert np.array_equal(got["window"], want["window"]) and np.array_equal(got["frame"], want["frame"])
    assert np.array_equal(got["svm_distance"], want["svm_distance"])
    for n in (1, 4, 6, 11):
        part = casc.detect(frames[:n], stage=capi.FDB_STAGE_NMS)
        ref = want[want["frame"] < n]
        assert np.array_equal(part["window"], ref["window"]) and np.array_equal(part["frame"], ref["frame"])
    capi.check(ctx.lib, ctx.lib.fdb_detect_batch(casc.h, padded.ctypes.data, 704, 0, capi.FDB_STAGE_NMS, None,
                                                 dets.ctypes.data, 4096, C.byref(cnt)))
    assert cnt.value == 0
    # error behaviour: too small result buffer, bad stage, pitch < width
    assert ctx.lib.fdb_detect_batch(casc.h, padded.ctypes.data, 704, 13, capi.FDB_STAGE_NMS, None, dets.ctypes.data, 1, C.byref(cnt)) == 6
    assert cnt.value == len(want)  # the needed capacity is reported
    assert ctx.lib.fdb_detect_batch(casc.h, padded.ctypes.data, 704, 1, 9, None, dets.ctypes.data, 4096, C.byref(cnt)) == 1
    assert ctx.lib.fdb_detect_batch(casc.h, padded.ctypes.data, 600, 1, 1, None, dets.ctypes.data, 4096, C.byref(cnt)) == 1


def test_model_validation_errors(ctx, face_models):
    """Loader-side errors of the reference become status codes (invalid_argument / unsupported)."""
    import copy
    import ctypes as C
    det_kw, wvm, svm = face_models
    bad = copy.copy(wvm)
    bad.rec = wvm.rec.copy(); bad.rec[0] = (0, 0, 25, 3)  # rectangle outside the 20x20 filter window
    d = bad.desc(); h = C.c_void_p()
    assert ctx.lib.fdb_wvm_create(ctx.h, C.byref(d), C.byref(h)) == 1 and not h.value
    with pytest.raises(capi.FdbError):  # patch size differs from the WVM filter size
        SlidingWindowCascade(ctx, dict(det_kw, patch_width=24), wvm, svm)
    with pytest.raises(capi.FdbError):  # DirectPyramidFeatureExtractor: stepX has to be greater than zero
        SlidingWindowCascade(ctx, dict(det_kw, step_x=-2), wvm, svm)
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    with pytest.raises(capi.FdbError):  # not prepared
        casc.detect_roi(syn.synthetic_frame(0), (0, 0, 10, 10))


def test_full_hd_frame(ctx, face_models):
    """BASELINE configs[2] geometry: one 1920x1080 frame (190 616 FaceFrontal windows, 21 pyramid layers up
    to 307 px wide -> many strips per layer), whole cascade against the oracle."""
    fo = _oracle()
    det_kw, wvm, svm = face_models
    frame = syn.synthetic_frame(5, 1920, 1080)
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(1920, 1080, 1)
    assert casc.windows_per_frame == 190616
    dets, dense = casc.detect(frame[None], stage=capi.FDB_STAGE_NMS, want_dense=True)
    ref = fo.detect_frame(det_kw, fo.Wvm(wvm), fo.Svm(svm), frame, stage=capi.FDB_STAGE_NMS)
    assert np.array_equal(dense[0]["level"], ref["dense"]["level"])
    assert np.max(np.abs(dense[0]["fout"] - ref["dense"]["fout"])) <= TOL
    assert list(dets["window"]) == list(ref["detections"]["window"])
    for idx in (casc.layers()[0]["index"], casc.layers()[-1]["index"]):
        _, layers = fo.pyramid(frame, det_kw["incremental_scale_factor"], det_kw["min_scale_factor"], det_kw["max_scale_factor"])
        img = [im for i, _, im in layers if i == idx][0]
        assert np.array_equal(casc.pyramid_layer(frame, idx), img)


def test_models_from_matlab_files(ctx, face_models, tmp_path):
    """the whole cascade with both classifiers read from MATLAB files (fdb_wvm_file_load / fdb_svm_mat_load): detections
    equal the oracle's for the same loaded numbers, and the cfg "threshold" applied afterwards behaves like the descriptor's"""
    pytest.importorskip("scipy.io")
    from featuredetection_b200.detector import load_wvm_mat, load_svm_mat
    from oracle import fdoracle as fo
    det_kw, wvm, svm = face_models
    c, t, sp, lg = str(tmp_path / "wvm.mat"), str(tmp_path / "thr.mat"), str(tmp_path / "svm.mat"), str(tmp_path / "log.mat")
    syn.write_wvm_mat(wvm, c, t, True)
    syn.write_svm_mat(svm, 20, 20, sp, lg)
    wl, sl = load_wvm_mat(c, t), load_svm_mat(sp, lg)
    assert np.array_equal(sl.sv, svm.sv) and np.array_equal(sl.coef, svm.coef)
    assert np.array_equal(wl.rec, wvm.rec) and np.array_equal(wl.hk_weights, wvm.hk_weights) and np.array_equal(wl.thresholds, wvm.thresholds)
    casc = SlidingWindowCascade(ctx, det_kw, wl, sl)
    casc.prepare(640, 480, 1)
    frame = syn.synthetic_frame(3)
    dets, dense = casc.detect(frame[None], stage=capi.FDB_STAGE_NMS, want_dense=True)
    ref = fo.detect_frame(det_kw, fo.Wvm(wl), fo.Svm(sl), frame, stage=capi.FDB_STAGE_NMS)
    assert np.array_equal(dense[0]["level"], ref["dense"]["level"])
    assert np.allclose(dense[0]["fout"], ref["dense"]["fout"], rtol=0, atol=TOL)
    assert list(dets["window"]) == list(ref["detections"]["window"])
    assert np.allclose(dets["svm_distance"], ref["detections"]["svm_distance"], rtol=0, atol=TOL)


def test_bgr_frames(ctx, face_models):
    """GrayscaleFilter's cvtColor branch on the device: gray frames bit-exact to the oracle's 2.4.3 formula (odd sizes, pitch,
    tails), and the cascade on BGR frames equals the cascade on the converted frames"""
    from featuredetection_b200.detector import gray_from_bgr
    from oracle import fdoracle as fo
    rng = np.random.default_rng(21)
    for (h, w, n) in ((480, 640, 3), (61, 97, 2), (5, 3, 1), (33, 1, 1)):
        bgr = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
        got = gray_from_bgr(ctx, bgr)
        for k in range(n):
            assert np.array_equal(got[k], fo.bgr_to_gray(bgr[k])), (h, w, k)
    padded = rng.integers(0, 256, (2, 50, 3 * 70 + 10), dtype=np.uint8)          # row pitch > 3 W
    view = np.stack([padded[k][:, :210].reshape(50, 70, 3) for k in range(2)])
    lib = ctx.lib
    out = np.empty((2, 50, 70), np.uint8)
    capi.check(lib, lib.fdb_gray_from_bgr(ctx.h, padded.ctypes.data, padded.shape[2], 70, 50, 2, out.ctypes.data))
    assert np.array_equal(out, np.stack([fo.bgr_to_gray(v) for v in view]))
    det_kw, wvm, svm = face_models
    casc = SlidingWindowCascade(ctx, det_kw, wvm, svm)
    casc.prepare(640, 480, 2)
    gray = syn.synthetic_frames(4, 2)
    colour = np.stack([np.clip(gray.astype(int) + d, 0, 255).astype(np.uint8) for d in (-9, 3, 5)], axis=-1)  # B, G, R
    conv = gray_from_bgr(ctx, colour)
    d1, dense1 = casc.detect_bgr(colour, want_dense=True)
    d2, dense2 = casc.detect(conv, want_dense=True)
    assert np.array_equal(dense1["level"], dense2["level"]) and np.array_equal(dense1["fout"], dense2["fout"])
    assert np.array_equal(d1["window"], d2["window"]) and np.array_equal(d1["svm_distance"], d2["svm_distance"])


@pytest.mark.parametrize("cntval", [2, 7, 12])
def test_any_number_of_grey_values_per_filter(ctx, cntval):
    """WvmClassifier::Area holds any number of grey values per filter (WvmClassifier.hpp:107-124); the synthetic benchmark
    models use 5. 2 (one rectangle value: the tensor-core strip path), 7 and 12 (more than the 4 columns per filter of the strip
    path: the generic kernels, 8 values per pass) against the oracle: dense records and stage-1 positives."""
    fo = _oracle()
    frame = syn.synthetic_frame(17)
    wvm = syn.make_wvm(20, 20, 6, 4, 0.04, seed=700 + cntval, cntval=cntval, rects_per_value=2 if cntval > 5 else 4)
    kw = dict(incremental_scale_factor=float(np.float32(0.92)), min_scale_factor=float(np.float32(0.05)),
              max_scale_factor=float(np.float32(0.16)), patch_width=20, patch_height=20, step_x=1, step_y=1,
              max_positives_per_frame=400000)
    # thresholds at the 35 % quantile of every third filter's output over the windows of another frame: exits at many levels
    probe = fo.detect_frame({k: v for k, v in kw.items() if k != "max_positives_per_frame"}, fo.Wvm(wvm), None, syn.synthetic_frame(16),
                            stage=capi.FDB_STAGE_WVM, want_patches=True, det_cap=400000)
    per_level = fo.Wvm(wvm).eval_all_levels(probe["patches"][::7])
    thr = np.full(24, -np.inf, np.float32)
    thr[1::3] = np.quantile(per_level[:, 1::3], 0.35, axis=0).astype(np.float32)
    wvm = wvm.with_thresholds(thr)
    casc = SlidingWindowCascade(ctx, kw, wvm, None)
    casc.prepare(640, 480, 2)
    frames = np.stack([frame, syn.synthetic_frame(18)])
    dets, dense = casc.detect(frames, stage=capi.FDB_STAGE_WVM, want_dense=True, det_cap=400000)
    okw = {k: v for k, v in kw.items() if k != "max_positives_per_frame"}
    for k in range(2):
        ref = fo.detect_frame(okw, fo.Wvm(wvm), None, frames[k], stage=capi.FDB_STAGE_WVM, det_cap=400000, frame_index=k)
        assert np.array_equal(dense[k]["level"], ref["dense"]["level"])
        assert np.max(np.abs(dense[k]["fout"] - ref["dense"]["fout"])) <= TOL
        assert list(dets[dets["frame"] == k]["window"]) == list(ref["detections"]["window"])
        assert len(np.unique(ref["dense"]["level"])) >= 4 and len(ref["detections"]) > 0
