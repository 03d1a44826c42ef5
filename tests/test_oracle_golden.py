"""The oracle (oracle/fd_oracle.c) against the committed golden vectors:
  tests/golden/cv2_primitives.npz   - OpenCV 4.13 outputs of cv::resize / cv::pyrDown
  tests/golden/ref_classifiers.npz  - outputs of the reference's OWN compiled sources (oracle/_ref)
and, when available in the running environment, against cv2 and oracle/_ref live."""
import os

import numpy as np
import pytest

from featuredetection_b200 import capi, synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fo(built):
    from oracle import fdoracle
    return fdoracle


@pytest.fixture(scope="module")
def cvg():
    return np.load(os.path.join(GOLD, "cv2_primitives.npz"))


@pytest.fixture(scope="module")
def refg():
    return np.load(os.path.join(GOLD, "ref_classifiers.npz"))


def test_resize_matches_cv2_golden(fo, cvg):
    for i, (sw, sh, dw, dh) in enumerate(cvg["resize_cases"]):
        got = fo.resize_linear(cvg["resize_src_%d" % i], int(dw), int(dh))
        assert np.array_equal(got, cvg["resize_dst_%d" % i]), (i, sw, sh, dw, dh)


def test_pyrdown_matches_cv2_golden(fo, cvg):
    for i in range(int(cvg["pyrdown_n"])):
        assert np.array_equal(fo.pyrdown(cvg["pyrdown_src_%d" % i]), cvg["pyrdown_dst_%d" % i]), i


def test_pyramid_chain_matches_cv2_golden(fo, cvg):
    f = cvg["chain_src"]
    for i in (1, 6):
        q = 0.5 ** (i / 8.0)
        a = fo.resize_linear(f, int(np.rint(320 * q)), int(np.rint(240 * q)))
        b = fo.pyrdown(a)
        c = fo.pyrdown(b)
        assert np.array_equal(a, cvg["chain_%d_0" % i]) and np.array_equal(b, cvg["chain_%d_1" % i]) and np.array_equal(c, cvg["chain_%d_2" % i])


def test_primitives_match_cv2_live(fo):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    rng = np.random.default_rng(9)
    for _ in range(25):
        sh, sw = int(rng.integers(3, 160)), int(rng.integers(3, 220))
        img = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        dh, dw = int(rng.integers(1, sh + 1)), int(rng.integers(1, sw + 1))
        assert np.array_equal(fo.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
        assert np.array_equal(fo.pyrdown(img), cv2.pyrDown(img))


def test_facefrontal_pyramid_table(fo):
    """layer indices, sizes and scales of the FaceFrontal pyramid at 640x480."""
    det_kw, _, _ = syn.landmark_models("FaceFrontal")
    olc, layers = fo.pyramid(syn.synthetic_frame(0), det_kw["incremental_scale_factor"], det_kw["min_scale_factor"], det_kw["max_scale_factor"])
    assert olc == 8
    assert [i for i, _, _ in layers] == list(range(22, 35))
    assert [im.shape[::-1] for _, _, im in layers] == [(96, 72), (88, 66), (80, 60), (74, 55), (68, 51), (62, 47), (57, 43),
                                                       (52, 39), (48, 36), (44, 33), (40, 30), (37, 28), (34, 26)]


def test_hq64_matches_reference_golden(fo, refg):
    for raw, eq in zip(refg["hq64_in"], refg["hq64_out"]):
        assert np.array_equal(fo.hq64(np.ascontiguousarray(raw)), eq)


@pytest.mark.parametrize("tag,profile", [("realistic", "realistic"), ("noexit", "no-exit")])
def test_wvm_matches_reference_golden(fo, refg, tag, profile):
    _, wvm, _ = syn.landmark_models("FaceFrontal", profile)
    lv, fout, pr, pos = fo.Wvm(wvm).eval(refg["patches"])
    assert np.array_equal(lv, refg["wvm_%s_level" % tag])
    assert np.array_equal(fout, refg["wvm_%s_fout" % tag])  # bit-exact float32
    assert np.array_equal(pr, refg["wvm_%s_prob" % tag]) and np.array_equal(pos, refg["wvm_%s_pos" % tag])


def test_svm_matches_reference_golden(fo, refg):
    _, _, svm = syn.landmark_models("FaceFrontal")
    d, p, q = fo.Svm(svm).eval(refg["patches"])
    assert np.array_equal(d, refg["svm_dist"]) and np.array_equal(p, refg["svm_prob"]) and np.array_equal(q, refg["svm_pos"])


@pytest.mark.parametrize("k", [0, 1])
def test_whole_frame_matches_reference_golden(fo, refg, k):
    det_kw, wvm, svm = syn.landmark_models("FaceFrontal")
    wo, so = fo.Wvm(wvm), fo.Svm(svm)
    frame = syn.synthetic_frame(k)
    r = fo.detect_frame(det_kw, wo, so, frame, stage=capi.FDB_STAGE_WVM)
    assert np.array_equal(r["dense"]["level"], refg["frame%d_dense_level" % k])
    assert np.array_equal(r["dense"]["fout"], refg["frame%d_dense_fout" % k])
    for stage, name in ((capi.FDB_STAGE_WVM, "wvm"), (capi.FDB_STAGE_OE, "oe"), (capi.FDB_STAGE_SVM, "svm"), (capi.FDB_STAGE_NMS, "nms")):
        got = fo.detect_frame(det_kw, wo, so, frame, stage=stage, want_dense=False)["detections"]["window"]
        assert list(got) == list(refg["frame%d_%s_windows" % (k, name)]), (k, name)


def test_overlap_elimination_matches_reference_golden(built, refg):
    import ctypes as C
    from featuredetection_b200.detector import DETECTION_DTYPE
    lib = capi.load_library()
    for i in range(4):
        cx, cy, w = refg["oe%d_in" % i]
        n = len(cx)
        d = np.zeros(n, DETECTION_DTYPE)
        d["center_x"], d["center_y"], d["width"], d["height"] = cx, cy, w, w
        d["probability"] = refg["oe%d_prob" % i]
        d["window"] = np.arange(n)
        dist, ratio = refg["oe%d_param" % i]
        cnt = C.c_int64()
        capi.check(lib, lib.fdb_overlap_eliminate(d.ctypes.data, n, float(dist), float(ratio), C.byref(cnt)))
        assert list(d[:cnt.value]["window"]) == list(refg["oe%d_keep" % i])


def test_oracle_matches_compiled_reference_live(fo):
    """When oracle/_ref is present: fresh random inputs, restatement vs the reference's own sources."""
    if not fo.ref_available():
        pytest.skip("oracle/_ref not built here")
    _, wvm, svm = syn.landmark_models("FaceFrontal")
    rng = np.random.default_rng(123)
    patches = rng.integers(0, 256, (150, 400), dtype=np.uint8)
    a, b = fo.Wvm(wvm).eval(patches), fo.Wvm(wvm, use_ref=True).eval(patches)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    a, b = fo.Svm(svm).eval(patches[:30]), fo.Svm(svm, use_ref=True).eval(patches[:30])
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    img = rng.integers(0, 256, (40, 50), dtype=np.uint8)
    assert np.array_equal(fo.hq64(img[3:27, 5:21]), fo.hq64(img[3:27, 5:21], use_ref=True))


def test_synthetic_inputs_are_reproducible():
    """The frame / model generators are pure integer or IEEE-exact numpy: pinned checksums prove that any
    machine (the GPU box included) regenerates the very inputs the golden vectors were made from."""
    import zlib
    assert zlib.crc32(syn.synthetic_frame(0).tobytes()) == 705085473
    assert zlib.crc32(syn.synthetic_frame(3).tobytes()) == 2467607839
    _, w, s = syn.landmark_models("FaceFrontal")
    assert zlib.crc32(w.val.tobytes()) == 2534942695
    assert zlib.crc32(w.hk_weights.tobytes()) == 3194455472
    assert zlib.crc32(s.sv.tobytes()) == 1226913186
    assert zlib.crc32(w.thresholds.tobytes()) == 1260564680


def test_svm_text_container_roundtrip(fo, tmp_path):
    """fdb_svm_file_load (product, host only) reads what the reference's own ProbabilisticSvmClassifier::store
    writes (needs oracle/_ref): float32 support vectors, coefficients, bias, gamma and the logistic line."""
    import ctypes as C
    if not fo.ref_available():
        pytest.skip("oracle/_ref not built here")
    lib = capi.load_library()
    rng = np.random.default_rng(4)
    model = syn.SvmModel(rng.normal(0, 1, (37, 144)).astype(np.float32), rng.normal(0, 1, 37).astype(np.float32),
                         gamma=0.2, bias=0.125, logistic_a=-0.6663, logistic_b=-1.8201)
    ref = fo.Svm(model, use_ref=True)
    path = str(tmp_path / "svm.txt").encode()
    assert fo.ref().ref_svm_store(ref.h, path) == 0
    h = C.c_void_p()
    capi.check(lib, lib.fdb_svm_file_load(path, C.byref(h)))
    d = lib.fdb_svm_file_desc(h).contents
    assert (d.num_sv, d.dim, d.sv_type, d.kernel) == (37, 144, capi.FDB_SV_F32, capi.FDB_KERNEL_RBF)
    # operator<< prints 6 significant digits: compare at that precision
    assert abs(d.gamma - 0.2) < 1e-6 and abs(d.bias - 0.125) < 1e-6
    assert abs(d.logistic_a + 0.6663) < 1e-6 and abs(d.logistic_b + 1.8201) < 1e-6
    coef = np.ctypeslib.as_array(d.coefficients, shape=(37,))
    sv = np.ctypeslib.as_array(C.cast(d.support_vectors, C.POINTER(C.c_float)), shape=(37, 144))
    assert np.allclose(coef, model.coef, rtol=1e-5, atol=1e-7) and np.allclose(sv, model.sv, rtol=1e-5, atol=1e-7)
    lib.fdb_svm_file_free(h)
    assert lib.fdb_svm_file_load(b"/nonexistent/file", C.byref(h)) == 2


def test_bgr_to_gray_formula(built):
    """GrayscaleFilter's cvtColor branch: the OpenCV 2.4.3 fixed-point formula; cv2 4.13 (15-bit coefficients) may differ by 1
    in a fraction of a percent of the pixels - the only executable check available for this primitive (parity unpinned)"""
    from oracle import fdoracle as fo
    rng = np.random.default_rng(77)
    bgr = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    got = fo.bgr_to_gray(bgr)
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    assert np.array_equal(got, ((1868 * b + 9617 * g + 4899 * r + 8192) >> 14).astype(np.uint8))
    for v in (0, 255):  # the coefficients sum to 2^14: grey stays grey
        assert np.all(fo.bgr_to_gray(np.full((4, 5, 3), v, np.uint8)) == v)
    cv2 = pytest.importorskip("cv2")
    ref = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 0.01


def test_reference_arm_with_cv2_pyramid_equals_the_restated_pyramid(fo):
    """bench.py's CPU reference arm builds its pyramids with cv2 (the SIMD cv::resize / cv::pyrDown the real reference links):
    same layers, same stage-1 records and detections as with the restated pyramid"""
    pytest.importorskip("cv2")
    if not fo.ref_available():
        pytest.skip("oracle/_ref not built")
    for name in ("FaceFrontal", "RightEyeCenter"):
        kw, wvm, svm = syn.landmark_models(name)
        frame = syn.synthetic_frame(11)[:240, :320] if name != "FaceFrontal" else syn.synthetic_frame(11)
        olc_a, a = fo.pyramid(frame, kw["incremental_scale_factor"], kw["min_scale_factor"], kw["max_scale_factor"])
        olc_b, b = fo.cv2_pyramid(frame, kw["incremental_scale_factor"], kw["min_scale_factor"], kw["max_scale_factor"])
        assert olc_a == olc_b and len(a) == len(b)
        for (ia, sa, xa), (ib, sb, xb) in zip(a, b):
            assert ia == ib and sa == sb and np.array_equal(xa, xb)
        rw, rs = fo.Wvm(wvm, use_ref=True), fo.Svm(svm, use_ref=True)
        r1 = fo.ref_detect_frame(kw, rw, rs, frame)
        r2 = fo.ref_detect_frame(kw, rw, rs, frame, pyramid_impl="cv2")
        assert r1["windows"] == r2["windows"] and np.array_equal(r1["dense"], r2["dense"])
        assert np.array_equal(r1["det_windows"], r2["det_windows"])
